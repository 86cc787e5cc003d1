"""CUDA-graph capture of one whole training step (forward + fused loss + backward + clamp/Adam).  The step
issues ~40 kernels whose total device time at the reference batch sizes (64-96) is under 1 ms: launched one by
one from Python the step is launch-bound, replayed from a graph it is not.  Everything step-dependent lives in
device memory (dropout counter, Adam step counters, learning rate -- re-synchronised from ``param_groups`` before every
replay), so a replay is a real, fresh training step.

Data parallel (world > 1): the step is captured as THREE graphs split at the two points where a gradient bucket
becomes final, and the NCCL all-reduces are issued eagerly between the replays:
    graph 0: forward + loss + vocab-projection backward   -> all-reduce bucket 0 (C.weight, C.bias), async
    graph 1: reverse-time recurrence + projection backward -> all-reduce bucket 1 (everything else), async
    (join the NCCL stream)   graph 2: fused clamp + Adam
so the first all-reduce still overlaps the reverse-time recurrence and no collective is ever captured.

Only fully teacher-forced steps are captured (teacher_forcing_ratio >= 1): scheduled sampling draws a host
coin per time step (stylenet/model.py:181) that changes the kernel sequence."""
import os

import torch

from . import ops
from .dp import merged_ranges


class GraphedTrainStep:
    def __init__(self, trainer, captions, lengths, features, warmup=3, force_segmented=False, **step_kw):
        """``trainer``: DataParallelTrainer (world size 1 or more).  ``captions`` / ``features``: example device
        tensors defining the static input shapes; ``lengths`` is fixed for the life of the graph."""
        if step_kw.get("teacher_forcing_ratio", 1.0) < 1.0:
            raise ValueError("only teacher-forced steps can be captured")
        self.trainer = trainer
        self.lengths = list(lengths)
        self.kw = step_kw
        self.captions = captions.clone()
        self.features = features.clone() if features is not None else None
        # warm up on a side stream (allocator, plans, lazily built state), then capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.trainer.step(self.captions, self.lengths, self.features, **self.kw)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.segments = []          # [(graph, ranges to all-reduce after it | None)]
        if (trainer.world == 1 or getattr(trainer, 'comm', 'nccl') == 'peer') and not force_segmented:
            g = torch.cuda.CUDAGraph()
            # SN_MAIN_PRIORITY=-1 captures on a high-priority stream (the main chain's kernel nodes inherit it and the
            # side streams yield the SMs to them).  Measured: the main chain then finishes ~10 us earlier, the weight-
            # gradient GEMMs and their Adam buckets ~10 us later -- same step time (profiles/README.md), so it is off
            prio = int(os.environ.get("SN_MAIN_PRIORITY", "0"))
            cap = torch.cuda.Stream(priority=prio) if prio != 0 else None
            with ops.no_gc_during_capture(), (torch.cuda.graph(g, stream=cap) if cap is not None else torch.cuda.graph(g)):
                self.loss, self.stats = self.trainer.step(self.captions, self.lengths, self.features, **self.kw)
            self.segments.append((g, None))
        else:
            self._capture_segmented()

    def _capture_segmented(self):
        tr = self.trainer
        dec = tr.decoder
        arena = dec.arena()
        state = {}

        def begin():
            g = torch.cuda.CUDAGraph()
            pool = self.segments[0][0].pool() if self.segments else None
            ctx = torch.cuda.graph(g, pool=pool) if pool is not None else torch.cuda.graph(g)
            ctx.__enter__()
            state["g"], state["ctx"] = g, ctx

        def end(ranges):
            state["ctx"].__exit__(None, None, None)
            self.segments.append((state["g"], ranges))

        def hook(names):               # a gradient bucket is final: close this graph, all-reduce eagerly at replay
            end(merged_ranges(arena, names))
            begin()

        with ops.no_gc_during_capture():
            begin()
            self.loss, self.stats = tr.forward_backward(self.captions, self.lengths, self.features, grad_hook=hook,
                                                        **self.kw)
            tr.optimizer.step()
            end(None)

    def __call__(self, captions=None, features=None):
        """Copy new inputs into the static buffers (async, same stream) and replay.  Returns the static loss
        tensor (valid once the stream reaches this point)."""
        if captions is not None:
            self.captions.copy_(captions, non_blocking=True)
        if features is not None:
            self.features.copy_(features, non_blocking=True)
        tr = self.trainer
        # the learning rate is read from device memory by the captured Adam kernels: refresh it (eagerly, same stream)
        # when adjust_learning_rate (stylenet/utils.py:116-124) changed param_groups since the last replay
        tr.optimizer._state()
        tr.optimizer._sync_lr()
        last = len(self.segments) - 1
        for i, (g, ranges) in enumerate(self.segments):
            if i == last and last > 0:
                tr.sync.wait()         # Adam waits for both all-reduces
            g.replay()
            tr.decoder.arena().kernel_epoch += 1      # the replay rewrote the parameters behind torch's back
            if ranges:
                tr.sync.launch(tr.decoder.arena().gflat, ranges)
        return self.loss, self.stats
