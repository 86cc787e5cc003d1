"""World-size-2 gloo test (CPU) of the data-parallel host logic: shard after the sort, bucketed SUM
all-reduce of flat arena ranges, token-weighted loss scaling -- the N>1 path minus the CUDA kernels."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import icei_b200 as sn
        torch.manual_seed(0)
        dec = sn.DecoderFactoredLSTM(12, 16, 20, 53, 1)
        a = dec.arena()
        lengths = [9, 8, 8, 6, 5, 3, 2]
        idx, ls = sn.shard_lengths(lengths, world, rank)
        n_global = sum(lengths)
        names_out = list(dec._out_names())
        names_seq = dec._seq_grad_names("sad")
        # fake "per-rank gradient scaled by 1/N_global": value = rank-specific constant * local tokens / N_global
        a.gflat.fill_(float(sum(ls)) / n_global)
        sync = sn.GradSync()
        sync.launch(a.gflat, sn.merged_ranges(a, names_out))       # bucket 0 (vocab projection)
        sync.launch(a.gflat, sn.merged_ranges(a, names_seq))       # bucket 1
        sync.wait()
        ok = True
        for n in names_out + names_seq:
            v = a.gflat[a.offset[n]:a.offset[n] + a.numel[n]]
            ok &= bool(torch.allclose(v, torch.ones_like(v)))       # sum_r N_r / N_global == 1
        # inactive styles were not reduced
        v = a.gflat[a.offset["S_happy_i.weight"]:a.offset["S_happy_i.weight"] + 4]
        ok &= bool(torch.allclose(v, torch.full_like(v, float(sum(ls)) / n_global)))
        q.put((rank, ok, sync.bytes, idx))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_bucketed_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert res[0][2] == res[1][2] > 0
    assert sorted(res[0][3] + res[1][3]) == list(range(7))
