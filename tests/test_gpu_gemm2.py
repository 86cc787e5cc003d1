"""CTA-pair tcgen05 GEMM (sn_gemm2.cu) and the fused vocabulary projection + log-softmax/NLL epilogues (GPU)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import icei_b200
    return icei_b200.ops


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _pad(v):
    return (v + 7) // 8 * 8


def _operands(op, M, N, K, g):
    if op == 2:
        A = torch.randn(K, _pad(M), device="cuda", generator=g).bfloat16()
    else:
        A = torch.randn(M, _pad(K), device="cuda", generator=g).bfloat16()
    if op == 0:
        B = torch.randn(N, _pad(K), device="cuda", generator=g).bfloat16()
    else:
        B = torch.randn(K, _pad(N), device="cuda", generator=g).bfloat16()
    a = A.double()[:, :K] if op != 2 else A.double()[:, :M].t()
    b = B.double()[:, :K].t() if op == 0 else B.double()[:, :N]
    return A, B, a @ b


@pytest.mark.parametrize("op", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (256, 256, 512), (96, 136, 72), (1920, 2048, 304), (257, 640, 1920),
                                   (2048, 512, 1920), (1920, 10000, 512), (5000, 300, 2048)])
def test_gemm_pair_bf16(ops, op, M, N, K):
    """Persistent cta_group::2 GEMM vs float64 matmul of the same bf16-rounded operands; more tiles than SM
    pairs (1920x10000: 320 tiles on 74 pairs) exercises the double-buffered accumulator hand-off."""
    g = torch.Generator(device="cuda").manual_seed(M + N * 3 + K * 7 + op)
    A, B, ab = _operands(op, M, N, K, g)
    bias = torch.randn(N, device="cuda", generator=g)
    C0 = torch.randn(M, N, device="cuda", generator=g)
    C = C0.clone()
    Cb = torch.zeros(M, _pad(N), device="cuda", dtype=torch.bfloat16)
    ops.gemm_bf16(op, A, B, M, N, K, A.stride(0), B.stride(0), C=C, ldc=N, Cb=Cb, ldcb=Cb.stride(0), bias=bias, beta=0.5,
                  impl="pair", splits=1 if K < 1024 else 0)
    want = ab + bias.double() + 0.5 * C0.double()
    assert _rel(C, want) < 1e-5
    assert _rel(Cb[:, :N].float(), want) < 6e-3


@pytest.mark.parametrize("op,M,N,K,splits", [(1, 1920, 512, 10000, 9), (2, 2048, 512, 1920, 3), (0, 300, 260, 4096, 16),
                                             (2, 10000, 512, 1920, 2)])
def test_gemm_pair_splitk(ops, op, M, N, K, splits):
    g = torch.Generator(device="cuda").manual_seed(M + N + K + op)
    A, B, ab = _operands(op, M, N, K, g)
    bias = torch.randn(N, device="cuda", generator=g)
    C = torch.full((M, N), 7.0, device="cuda")
    ops.gemm_bf16(op, A, B, M, N, K, A.stride(0), B.stride(0), C=C, ldc=N, bias=bias, impl="pair", splits=splits)
    assert _rel(C, ab + bias.double()) < 1e-5
    C2 = torch.full((M, N), 7.0, device="cuda")
    ops.gemm_bf16(op, A, B, M, N, K, A.stride(0), B.stride(0), C=C2, ldc=N, bias=bias, impl="pair", splits=splits)
    assert torch.equal(C, C2)            # work-space reduction is deterministic


def test_gemm_pair_grouped(ops):
    n, F, H = 700, 512, 512
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(n, 4 * F, device="cuda", generator=g).bfloat16()
    B = (torch.randn(4, H, F, device="cuda", generator=g) / 16).bfloat16()
    bias = torch.randn(4 * H, device="cuda", generator=g)
    C = torch.zeros(n, 4 * H, device="cuda")
    ops.gemm_bf16(0, A, B, n, H, F, 4 * F, F, C=C, ldc=4 * H, bias=bias, batch=4, sA=F, sB=H * F, sC=H, sBias=H, impl="pair")
    want = torch.cat([A[:, i * F:(i + 1) * F].double() @ B[i].double().t() + bias[i * H:(i + 1) * H].double()
                      for i in range(4)], 1)
    assert _rel(C, want) < 1e-5


@pytest.mark.parametrize("N,V,H", [(1920, 10000, 512), (77, 1003, 64), (300, 257, 520), (8192, 10000, 512)])
def test_vocab_fused_nll(ops, N, V, H):
    """Fused projection + log-softmax/NLL: statistics against torch on the logits the same GEMM kernel stores
    (bit-identical accumulation -> argmax / rank exact), loss and gradient against float64."""
    g = torch.Generator(device="cuda").manual_seed(N + V + H)
    Hb = torch.randn(N, _pad(H), device="cuda", generator=g).bfloat16()
    Wb = (torch.randn(V, _pad(H), device="cuda", generator=g) * (3.0 / H ** 0.5)).bfloat16()
    if _pad(H) != H:
        Hb[:, H:] = 0
        Wb[:, H:] = 0
    bias = torch.randn(V, device="cuda", generator=g)
    tgt = torch.randint(0, V, (N,), device="cuda", generator=g)
    logits = torch.empty(N, V, device="cuda")
    ops.gemm_bf16(0, Hb, Wb, N, V, H, Hb.stride(0), Wb.stride(0), C=logits, ldc=V, bias=bias, impl="pair")
    f32 = dict(dtype=torch.float32, device="cuda")
    tl, lse, row_loss = torch.empty(N, **f32), torch.empty(N, **f32), torch.empty(N, **f32)
    am = torch.empty(N, dtype=torch.int64, device="cuda")
    above = torch.full((N,), 123, dtype=torch.int32, device="cuda")
    ops.vocab_nll_fwd(Hb, Wb, bias, tgt, N, V, H, tl, lse, row_loss=row_loss, argmax=am, above=above)
    assert torch.equal(am, logits.argmax(1))
    assert torch.equal(tl, logits.gather(1, tgt[:, None])[:, 0])
    want_lse = torch.logsumexp(logits.double(), 1)
    assert (lse.double() - want_lse).abs().max().item() < 2e-5
    assert _rel(row_loss, want_lse - logits.double().gather(1, tgt[:, None])[:, 0]) < 1e-5
    assert int(above.abs().sum()) == 0
    Vp = _pad(V)
    dLb = torch.full((N, Vp), 9.0, device="cuda", dtype=torch.bfloat16)
    t5 = torch.empty(N, dtype=torch.int32, device="cuda")
    scale = 1.0 / N
    ops.vocab_nll_bwd(Hb, Wb, bias, tgt, N, V, H, tl, lse, scale, dLb=dLb, above=above, top5hit=t5)
    want_above = (logits > tl[:, None]).sum(1).to(torch.int32)
    assert torch.equal(above, want_above)
    assert torch.equal(t5, (want_above < 5).to(torch.int32))
    p = torch.softmax(logits.double(), 1)
    p[torch.arange(N, device="cuda"), tgt] -= 1.0
    assert _rel(dLb[:, :V].float(), p * scale) < 6e-3
    if Vp != V:
        assert float(dLb[:, V:].float().abs().max()) == 0.0
    # ranking-only pass (no gradient written)
    above.zero_()
    ops.vocab_nll_bwd(Hb, Wb, bias, tgt, N, V, H, tl, lse, scale, dLb=None, above=above, top5hit=t5)
    assert torch.equal(above, want_above)


@pytest.mark.parametrize("cell", [0, 1])
@pytest.mark.parametrize("B,H,lengths", [(1100, 512, None), (700, 64, "ragged"), (1300, 256, "ragged")])
def test_recurrence_gemm_form(ops, cell, B, H, lengths):
    """Large-batch K3 (one CTA-pair GEMM per step, cell fused into the epilogue): forward against a float64 unroll with
    the same bf16 roundings of W_hh / h, backward against the fp32 persistent kernel on the same saved activations."""
    import icei_b200
    T = 7
    g = torch.Generator().manual_seed(B * 7 + H + cell + 100)
    if lengths is None:
        lengths = [T] * B
    else:
        lengths = sorted(torch.randint(2, T + 1, (B,), generator=g).tolist(), reverse=True)
        lengths[0] = T
    plan = icei_b200.get_plan(lengths)
    d = plan.dev("cuda")
    N, T = plan.N, plan.T
    XP = (torch.randn(N, 4 * H, generator=g) * 0.7).cuda()
    W = (torch.randn(4 * H, H, generator=g) / H ** 0.5).cuda()
    Wb = W.bfloat16().contiguous()
    Wil = ops.cast_gate_interleave(W)
    # the interleaved shadow is a row permutation of the plain one
    n = torch.arange(4 * H, device="cuda")
    src = ((n >> 6) & 3) * H + (n >> 8) * 64 + (n & 63)
    assert torch.equal(Wil, Wb[src])
    bhh = (torch.randn(4 * H, generator=g) * 0.1).cuda()
    dH = torch.randn(N, H, generator=g).cuda()
    Hall = torch.empty(N, H, device="cuda"); Call = torch.empty(N, H, device="cuda")
    Hb = torch.empty(N, H, device="cuda", dtype=torch.bfloat16); Hpb = torch.empty_like(Hb)
    gates = torch.empty(N, 4 * H, device="cuda")
    ops.recur_fwd_gemm(cell, H, B, plan, XP, Wil, bhh, Hall, Hb, Call, gates)
    ops.recur_hprev(Hb, d, N, H, Hpb)
    torch.cuda.synchronize()
    w, b, xp = Wb.double().cpu(), bhh.double().cpu(), XP.double().cpu()
    h = torch.zeros(B, H, dtype=torch.float64); c = torch.zeros(B, H, dtype=torch.float64)
    hs, cs, hp = [], [], []
    for t, bt in enumerate(plan.bs):
        hq = h[:bt].bfloat16().double()
        hp.append(hq)
        z = xp[plan.off[t]:plan.off[t] + bt] + hq @ w.t() + b
        if cell == 0:
            i, f, o, gg = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        else:
            i, f, gg, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        c = torch.sigmoid(f) * c[:bt] + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * (c if cell == 0 else torch.tanh(c))
        hs.append(h); cs.append(c)
    hall, call = torch.cat(hs, 0), torch.cat(cs, 0)
    assert _rel(Hall.cpu(), hall) < 3e-3
    assert _rel(Call.cpu(), call) < 3e-3
    assert _rel(Hb.float().cpu(), hall) < 6e-3
    assert _rel(Hpb.float().cpu(), torch.cat(hp, 0)) < 6e-3
    # backward vs the fp32 persistent kernel on the SAME saved activations (only the dZ exchange is rounded)
    dZ = torch.empty(N, 4 * H, device="cuda"); dZb = torch.empty(N, 4 * H, device="cuda", dtype=torch.bfloat16)
    dc = torch.empty(B, H, device="cuda")
    ops.recur_bwd_gemm(cell, H, B, plan, Wb, Call, gates, dH, dZ, dZb, dc)
    dZ32 = torch.empty(N, 4 * H, device="cuda")
    dh2 = torch.zeros(B, H, device="cuda"); dc2 = torch.zeros(B, H, device="cuda")
    ops.recur_bwd(cell, H, B, d["bs"], d["off"], 0, T, Wb.float(), None, Call, gates, dH, dZ32, dh2, dc2)
    assert _rel(dZ, dZ32) < 1e-2
    assert _rel(dZb.float(), dZ32) < 1.5e-2
    assert _rel(dc, dc2) < 1e-2


def test_large_batch_training_step_uses_gemm_recurrence_and_matches():
    """End to end at a batch above ops.RECUR_GEMM_MIN_BATCH: loss and gradients of the bf16 decoder with the per-step
    GEMM recurrence == the same decoder forced onto the persistent kernel (both bf16 mode), within bf16 noise."""
    import random
    import icei_b200 as sn
    from icei_b200 import ops
    from oracle import port
    E, H, F, V, B, T = 300, 512, 512, 1000, 1200, 6
    torch.manual_seed(3)
    dec = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0).cuda().set_precision("bf16")
    dec.train()
    cap, lens, feats = port.synthetic_batch(B, T, V, E=E, ragged=True, seed=4)
    cap, feats = cap.cuda(), feats.cuda()
    res = []
    for thr in (10 ** 9, 1024):
        ops.RECUR_GEMM_MIN_BATCH[0] = thr
        dec.zero_grad()
        random.seed(1)
        loss, _ = dec.forward_loss(cap, lens, feats, teacher_forcing_ratio=1.0, mode="happy")
        res.append((loss.item(), {n: p.grad.clone() for n, p in dec.named_parameters() if p.grad is not None}))
    ops.RECUR_GEMM_MIN_BATCH[0] = 1024
    assert abs(res[0][0] - res[1][0]) < 2e-3 * abs(res[0][0])
    assert set(res[0][1]) == set(res[1][1])
    for n in res[0][1]:
        assert _rel(res[1][1][n], res[0][1][n]) < 2e-2, n
