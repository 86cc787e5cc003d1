#!/usr/bin/env python
"""Decode throughput (BASELINE.json metrics 2 and 3): greedy-decode captions/s and beam-search (k=5)
captions/s swept over images per call, on one GPU, next to the CPU oracle port on the host cores.

    python bench_decode.py [--precision fp32|bf16] [--max-images 1024] [--cpu-images 8]
    torchrun --nproc-per-node N bench_decode.py ...     # N replicas (decode shards images, no collective): every rank
                                                        # decodes its own `images_per_call`, rank 0 prints the SUM

Greedy = the reference's validation path ``forward(teacher_forcing_ratio=0)`` under no_grad
(stylenet/train_multitask.py:296-299).  Beam = ``sample(k=5)`` semantics (stylenet/model.py:198-294 with the
app/backend image feed), batched over images here, one image per call in the reference
(stylenet/evaluator.py:74-81).  Weights: reference init + the decode recipe of SURVEY.md section 8c
(C.weight x30, C.bias[<end>] = +2) so captions terminate at varied lengths.  One JSON line per row."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
V, E, H, F, T = 10000, 300, 512, 512, 20


def sharpen(dec):
    import torch
    out = dec.C if hasattr(dec, "C") else dec.linear
    with torch.no_grad():
        out.weight.mul_(30.0)
        out.bias[2] = 2.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--max-images", type=int, default=1024)
    ap.add_argument("--cpu-images", type=int, default=8)
    args = ap.parse_args()
    import torch
    import icei_b200 as sn
    from icei_b200.decode import beam_sample
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    dec = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.5)
    sharpen(dec)
    sd = {k: v.clone() for k, v in dec.state_dict().items()}
    dec = dec.to(dev).eval().set_precision(args.precision)
    g = torch.Generator().manual_seed(1 + rank)
    out_rows = []

    def timed(fn, reps):
        for _ in range(3):          # warm-up: the 2nd call of a decode session captures its CUDA graph
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    # ---- greedy: forward(tf=0), batch 96, T=20 -------------------------------------------------------
    B = 96
    cap = torch.randint(4, V, (B, T), generator=g)
    cap[:, 0] = 1
    feats = torch.randn(B, E, generator=g)
    cap_d, feats_d = cap.to(dev), feats.to(dev)
    with torch.no_grad():
        dt = timed(lambda: dec(cap_d, [T] * B, feats_d, teacher_forcing_ratio=0.0, mode="happy"), 10)
    out_rows.append({"metric": "greedy_decode_captions_per_sec", "impl": "b200", "precision": args.precision,
                     "value": B / dt, "unit": "captions/s", "batch": B, "steps": T,
                     "path": "forward(teacher_forcing_ratio=0) under no_grad"})
    # ---- beam sweep -------------------------------------------------------------------------------------
    n = 1
    lens_seen = None
    while n <= args.max_images:
        f = torch.randn(n, E, generator=g).to(dev)
        for k in (1, 5):
            reps = 5 if n <= 64 else 2
            res = []
            dt = timed(lambda: res.append(beam_sample(dec, f, 1, 2, k, "happy", True)), reps)
            lens = [r.shape[1] for r in res[-1]]
            out_rows.append({"metric": "beam_decode_captions_per_sec", "impl": "b200", "precision": args.precision,
                             "k": k, "images_per_call": n, "value": n / dt, "unit": "captions/s",
                             "mean_caption_len": sum(lens) / len(lens), "ms_per_call": dt * 1e3})
        n *= 4
    if world > 1:
        # replicas only: aggregate = sum over ranks of each rank's own throughput on its own images
        import torch.distributed as dist
        vals = torch.tensor([r["value"] for r in out_rows], dtype=torch.float64, device=dev)
        lo = vals.clone()
        dist.all_reduce(vals, op=dist.ReduceOp.SUM)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        for r, v, m in zip(out_rows, vals.tolist(), lo.tolist()):
            r.update(value=v, n_gpus=world, slowest_rank_value=m, images_per_call_total=r.get("images_per_call", r.get("batch")) * world)
        if rank == 0:
            for r in out_rows:
                print(json.dumps(r), flush=True)
        dist.destroy_process_group()
        return
    # ---- CPU oracle port on the host cores -----------------------------------------------------------
    from oracle import port
    torch.set_num_threads(os.cpu_count() or 1)
    ref = port.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.5)
    ref.load_state_dict(sd)
    ref.eval()
    with torch.no_grad():
        ref(cap, [T] * B, feats, teacher_forcing_ratio=0.0, mode="happy")
        t0 = time.perf_counter()
        for _ in range(3):
            ref(cap, [T] * B, feats, teacher_forcing_ratio=0.0, mode="happy")
        dt = (time.perf_counter() - t0) / 3
    out_rows.append({"metric": "greedy_decode_captions_per_sec", "impl": "reference(port,cpu)", "value": B / dt,
                     "unit": "captions/s", "batch": B, "cores": os.cpu_count()})
    fcpu = torch.randn(args.cpu_images, E, generator=g)
    for k in (1, 5):
        with torch.no_grad():
            t0 = time.perf_counter()
            lens = [ref.sample(fcpu[i:i + 1], 1, 2, k=k, mode="happy", feed_image=True).shape[1]
                    for i in range(args.cpu_images)]
            dt = time.perf_counter() - t0
        out_rows.append({"metric": "beam_decode_captions_per_sec", "impl": "reference(port,cpu)", "k": k,
                         "images_per_call": 1, "value": args.cpu_images / dt, "unit": "captions/s",
                         "mean_caption_len": sum(lens) / len(lens), "cores": os.cpu_count()})
    for r in out_rows:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
