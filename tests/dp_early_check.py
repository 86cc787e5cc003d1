"""Run under torchrun (one rank per GPU), bf16 mode, peer-fused exchange: issuing the exchange+Adam of the vocabulary
projection EARLY (side stream, under the reverse recurrence; dp.EARLY_PEER_EXCHANGE) gives the same parameters as the
single exchange at the end of the step.  Prints 'DP_EARLY_OK' on rank 0.  Launched by tests/test_gpu_dp.py."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import icei_b200 as sn
    from icei_b200 import dp
    from oracle import port
    V, E, H, F, B, T = 517, 40, 64, 64, 64, 9
    cap, lens, feats = port.synthetic_batch(B, T, V, E=E, ragged=True, seed=3)
    n_global = sum(lens)
    idx, my_lens = sn.shard_lengths(lens, world, rank)
    cap_d, feats_d = cap[idx].to(dev), feats[idx].to(dev)
    res = []
    # all early / all late / ALTERNATING: the set of ranges in one exchange call changes from step to step in the last
    # schedule, which must not move the ownership of an element (its Adam moments live on exactly one rank)
    for schedule in ((True,) * 4, (False,) * 4, (True, False, True, False)):
        torch.manual_seed(0)
        dec = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0).to(dev).train().set_precision("bf16")
        tr = sn.DataParallelTrainer(dec, sn.FusedClampAdam(dec, lr=5e-4), comm="peer")
        for early in schedule:
            dp.EARLY_PEER_EXCHANGE[0] = early
            loss, _ = tr.step(cap_d, my_lens, feats_d, n_global=n_global, mode="sad")
        torch.cuda.synchronize()
        dist.barrier()
        res.append((loss.item(), {n: p.detach().clone() for n, p in dec.named_parameters()}))
    dp.EARLY_PEER_EXCHANGE[0] = True
    worst = max(((res[k][1][n] - res[1][1][n]).norm() / res[1][1][n].norm().clamp_min(1e-30)).item()
                for k in (0, 2) for n in res[0][1])
    ok = worst < 1e-5 and all(abs(res[k][0] - res[1][0]) < 1e-5 * abs(res[1][0]) for k in (0, 2))
    # every rank must also hold the same parameters as rank 0 (the all-gather of both calls reached everybody)
    for n, p in res[0][1].items():
        q = p.clone()
        dist.broadcast(q, 0)
        ok &= bool(torch.equal(p, q))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_EARLY_OK" if flag.item() == 1.0 else "DP_EARLY_FAIL", "worst param rel diff %.2e" % worst, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
