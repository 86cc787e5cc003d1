"""Batch-sharded data parallelism over the GPUs of one box: one process per GPU,
weights / Adam states / teacher-forcing coins replicated, samples sharded after the length sort, and ONE
exchange step -- a SUM all-reduce (NCCL over NVLink/NVSwitch) of the flat gradient-arena ranges that
received a gradient this step, issued in two buckets as backward produces them so the first (the
vocabulary projection, 1/3 of the bytes) overlaps the reverse-time recurrence:
    bucket 0: C.weight, C.bias          ready right after the vocab-projection backward
    bucket 1: everything else active    ready after the sequence backward
Each rank scales its gradients by 1/N_global tokens (not 1/N_local) so the SUM equals the single-GPU
gradient of the token-mean loss even with ragged shards.  The reference has no distributed code at all.
"""
import torch
import torch.distributed as dist


def merged_ranges(arena, names, max_gap=64):
    """Flat (offset, length) ranges covering ``names`` in the arena, merged across alignment padding."""
    spans = sorted((arena.offset[n], arena.numel[n]) for n in names)
    out = []
    for o, n in spans:
        if out and o - (out[-1][0] + out[-1][1]) < max_gap:
            out[-1] = (out[-1][0], o + n - out[-1][0])
        else:
            out.append((o, n))
    return out


class GradSync:
    """Asynchronous SUM all-reduce of flat ranges of one buffer; ``wait()`` joins the current stream."""

    def __init__(self, group=None):
        self.group = group
        self.pending = []
        self.bytes = 0

    def launch(self, flat, ranges):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return                      # single process: nothing to exchange
        for o, n in ranges:
            w = dist.all_reduce(flat[o:o + n], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.pending.append(w)
            self.bytes += 4 * n

    def wait(self):
        for w in self.pending:
            w.wait()
        self.pending = []


class PeerExchange:
    """NVLink peer-memory view of every rank's parameter arena, gradient arena and signal pad: each rank exports
    CUDA IPC handles of the three buffers (exchanged once through torch.distributed), and opens its peers' with
    its own device current, which yields peer mappings its kernels can load from / store to over NVLink.
    Feeds sn_dp_adam_fused: ONE kernel per step does reduce-scatter + clamp/Adam + all-gather, no NCCL."""

    _opened = {}        # handle bytes -> mapped base address (an allocation can be opened once per process)
    N_PADS = 16         # exchange calls of one step that may be in flight at the same time (one pad each)

    def pads(self, bucket):
        """Signal-pad pointers (one per rank) of bucket ``bucket``: concurrent exchange kernels on different streams
        must not share flags."""
        if not 0 <= bucket < self.N_PADS:
            raise RuntimeError("too many exchange buckets in one step")
        return [p + 128 * bucket for p in self.pad_ptrs]

    def __init__(self, arena, group=None):
        from . import ops
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("peer exchange supports one NVSwitch box (<= 8 ranks)")
        dev = arena.flat.device
        self.pad = torch.zeros(32 * self.N_PADS, dtype=torch.int32, device=dev)   # one signal pad per concurrent bucket
        torch.cuda.synchronize()
        mine = [ops.ipc_export(t) for t in (arena.flat, arena.gflat, self.pad)]
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (self.rank, dev.index, mine), group=group)
        self.param_ptrs, self.grad_ptrs, self.pad_ptrs = [0] * self.world, [0] * self.world, [0] * self.world
        for r, dev_index, exports in gathered:
            if r == self.rank:
                ptrs = [t.data_ptr() for t in (arena.flat, arena.gflat, self.pad)]
            else:
                ops.enable_peer_access(dev_index)
                ptrs = []
                for handle, offset in exports:
                    base = PeerExchange._opened.get(handle)
                    if base is None:
                        base = ops.ipc_open(handle)
                        PeerExchange._opened[handle] = base
                    ptrs.append(base + offset)
            self.param_ptrs[r], self.grad_ptrs[r], self.pad_ptrs[r] = ptrs
        self.arena_version = arena.version
        self.keep = (arena.flat, arena.gflat, self.pad)      # exported memory must stay allocated
        torch.cuda.synchronize()
        dist.barrier(group=group)


EARLY_PEER_EXCHANGE = [True]     # split the peer-fused exchange+Adam: vocabulary projection early, the rest at the end
PEER_BUCKETS = [False]           # also exchange the W_hh/U, S/V and embedding buckets as soon as they are final (slower: see step())


class DataParallelTrainer:
    """forward -> loss -> backward -> gradient exchange -> fused clamp+Adam, one rank per GPU.

    comm="peer" (default when world > 1 on CUDA): the exchange and the optimizer are ONE kernel over NVLink peer
    memory (sn_dp_adam_fused: reduce-scatter by peer loads, Adam on the owned shard, all-gather by peer stores).
    comm="nccl": bucketed NCCL SUM all-reduce (bucket 0 overlapped with the recurrence backward) + local Adam."""

    def __init__(self, decoder, optimizer, group=None, comm=None):
        self.decoder, self.optimizer = decoder, optimizer
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.sync = GradSync(group)
        if comm is None:
            comm = "peer" if (self.world > 1 and dist.get_backend(group) == "nccl") else "nccl"
        self.comm = comm
        self.peers = None

    def _peers(self):
        a = self.decoder.arena()
        if self.peers is None or self.peers.arena_version != a.version:
            self.peers = PeerExchange(a, self.group)
        return self.peers

    def forward_backward(self, captions, lengths, features, n_global=None, b_global=None, grad_hook=None, **kw):
        """zero_grad + forward + loss + backward with gradients scaled by 1/N_global; ``grad_hook(names)`` is
        called when a bucket of gradients is final."""
        dec = self.decoder
        a = dec.arena()
        N_local = sum(int(l) for l in lengths)
        if n_global is None:
            n_global = N_local * self.world
        for p in a.named.values():
            p.grad = None
        extra = {}
        if b_global is not None or hasattr(dec, "attention"):
            extra["b_global"] = b_global if b_global is not None else len(lengths) * self.world
        return dec.forward_loss(captions, lengths, features, n_global=n_global, grad_hook=grad_hook, **extra, **kw)

    def step(self, captions, lengths, features, n_global=None, b_global=None, **kw):
        a = self.decoder.arena()
        bucketed = getattr(self.decoder, "bf16", False) and not hasattr(self.decoder, "attention") and EARLY_PEER_EXCHANGE[0]
        if self.world > 1 and self.comm == "peer":
            peers = self._peers()
            early = []
            if bucketed:
                # gradient buckets are exchanged (reduce-scatter + Adam + all-gather, one kernel) on the side stream that
                # produced them AS SOON AS THEY ARE FINAL: the vocabulary projection under the reverse recurrence, W_hh / U
                # and S / V under the rest of the projection backward; only the embedding is left for the end of the step.
                # Every bucket has its own signal pad (the kernels overlap each other).
                calls = [0]
                out_names = set(self.decoder._out_names())

                def early_step(names):
                    # every exchange kernel costs two cross-GPU barriers plus the skew between the ranks (~50 us measured
                    # at N=2, profiles/r2_timeline_N2.txt): only the large vocabulary bucket, which has the whole reverse
                    # recurrence + projection backward to hide under, goes early; everything else is ONE call at the end
                    if not PEER_BUCKETS[0] and not out_names.issuperset(names):
                        return
                    self.optimizer.step_peer(peers, only=names, bucket=calls[0])
                    calls[0] += 1
                    early.extend(names)
                kw = dict(kw, early_step=early_step)
            loss, stats = self.forward_backward(captions, lengths, features, n_global=n_global, b_global=b_global, **kw)
            self.optimizer.step_peer(peers, skip=early or None, bucket=PeerExchange.N_PADS - 1)
            return loss, stats
        hook = None
        if self.world > 1:
            def hook(names):
                self.sync.launch(a.gflat, merged_ranges(a, names))
        early = []
        if self.world == 1 and bucketed:
            # single GPU: Adam of each bucket runs on the side stream that produced it (under the reverse recurrence /
            # the projection backward); the rest follows at the end
            def early_step(names):
                self.optimizer.step(only=names)
                early.extend(names)
            kw = dict(kw, early_step=early_step)
        loss, stats = self.forward_backward(captions, lengths, features, n_global=n_global, b_global=b_global,
                                            grad_hook=hook, **kw)
        if self.world > 1:
            self.sync.wait()
        self.optimizer.step(skip=early or None)
        return loss, stats
