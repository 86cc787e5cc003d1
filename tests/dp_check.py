"""Run under torchrun (one rank per GPU): data-parallel step == single-GPU step on the same global batch.
Prints 'DP_CHECK_OK' on rank 0.  Launched by tests/test_gpu_dp.py."""
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
    from oracle import port
    V, E, H, F, B, T = 500, 28, 64, 72, 16, 9
    torch.manual_seed(0)
    make = lambda: sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0).to(dev).train()
    dec, ref = make(), make()
    ref.load_state_dict(dec.state_dict())
    cap, lens, feats = port.synthetic_batch(B, T, V, E=E, ragged=True, seed=3)
    n_global = sum(lens)
    # single GPU, full batch
    ref.zero_grad()
    loss_ref, _ = ref.forward_loss(cap.to(dev), lens, feats.to(dev), mode="sad")
    opt_ref = sn.FusedClampAdam(ref, lr=5e-4)
    g_ref = {n: p.grad.clone() for n, p in ref.named_parameters() if p.grad is not None}
    opt_ref.step()
    # data parallel
    idx, my_lens = sn.shard_lengths(lens, world, rank)
    opt = sn.FusedClampAdam(dec, lr=5e-4)
    comm = os.environ.get("SN_DP_COMM", "peer")
    tr = sn.DataParallelTrainer(dec, opt, comm=comm)
    g_local = None
    loss, _ = tr.step(cap[idx].to(dev), my_lens, feats[idx].to(dev), n_global=n_global, mode="sad")
    if comm == "peer":
        # the fused kernel leaves the LOCAL gradients in .grad: reduce them here only to compare
        for n, p in dec.named_parameters():
            if p.grad is not None:
                g = p.grad.clone()
                dist.all_reduce(g)
                p.grad = g
    tot = loss.clone()
    dist.all_reduce(tot)
    torch.cuda.synchronize()
    ok = abs(tot.item() - loss_ref.item()) < 1e-5 * abs(loss_ref.item())
    worst = 0.0
    for n, p in dec.named_parameters():
        if n in g_ref:
            e = ((p.grad - g_ref[n]).norm() / g_ref[n].norm().clamp_min(1e-30)).item()
            worst = max(worst, e)
        else:
            ok &= p.grad is None
    ok &= worst < 1e-5
    for (n, p), (_, q) in zip(dec.named_parameters(), ref.named_parameters()):
        ok &= bool(((p - q).norm() / q.norm().clamp_min(1e-30)).item() < 1e-5)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_CHECK_OK" if flag.item() == 1.0 else "DP_CHECK_FAIL", comm, "worst grad rel err %.2e" % worst,
              "loss %.6f vs %.6f" % (tot.item(), loss_ref.item()), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
