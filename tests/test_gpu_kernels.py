"""Kernel-level parity (GPU): each C-ABI entry point against a few lines of torch on the same inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import icei_b200
    return icei_b200.ops


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("op", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 7, 3), (96, 130, 300), (257, 64, 129), (1920, 512, 300)])
def test_gemm_simt(ops, op, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 1000 + N * 10 + K + op)
    A = torch.randn((M, K) if op != 2 else (K, M), device="cuda", generator=g)
    B = torch.randn((N, K) if op == 0 else (K, N), device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    C0 = torch.randn(M, N, device="cuda", generator=g)
    C = C0.clone()
    ops.gemm(op, A, B, C, M, N, K, A.stride(0), B.stride(0), N, bias=bias, beta=0.5)
    a = A.double() if op != 2 else A.double().t()
    b = B.double().t() if op == 0 else B.double()
    want = a @ b + bias.double() + 0.5 * C0.double()
    assert _rel(C, want) < 3e-6


def test_gemm_batched_strided(ops):
    # the 4-gate grouped layout: A [n,4F] column blocks, B [4][H,F], C [n,4H] column blocks
    n, F, H = 37, 20, 16
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(n, 4 * F, device="cuda", generator=g)
    B = torch.randn(4, H, F, device="cuda", generator=g)
    bias = torch.randn(4 * H, device="cuda", generator=g)
    C = torch.zeros(n, 4 * H, device="cuda")
    ops.gemm(0, A, B, C, n, H, F, 4 * F, F, 4 * H, bias=bias, batch=4, sA=F, sB=H * F, sC=H, sBias=H)
    want = torch.cat([A[:, i * F:(i + 1) * F].double() @ B[i].double().t() + bias[i * H:(i + 1) * H].double()
                      for i in range(4)], 1)
    assert _rel(C, want) < 3e-6


def test_colsum(ops):
    X = torch.randn(1000, 77, device="cuda")
    out = torch.ones(77, device="cuda")
    ops.colsum(X, 1000, 77, 77, out, beta=2.0)
    assert _rel(out, X.double().sum(0) + 2.0) < 1e-6


@pytest.mark.parametrize("V", [53, 10000])
def test_softmax_nll(ops, V):
    N = 33
    g = torch.Generator(device="cuda").manual_seed(V)
    x = torch.randn(N, V, device="cuda", generator=g) * 3
    x[3, 7] = x[3].max() + 1.0
    x[3, 5] = x[3, 7]                      # exact tie -> lowest index wins
    t = torch.randint(0, V, (N,), device="cuda", generator=g)
    row = torch.empty(N, device="cuda")
    d = torch.empty_like(x)
    am = torch.empty(N, dtype=torch.int64, device="cuda")
    t5 = torch.empty(N, dtype=torch.int32, device="cuda")
    ops.softmax_nll(x, N, V, targets=t, row_loss=row, dlogits=d, grad_scale=1.0 / N, argmax=am, top5hit=t5)
    xd = x.double().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(xd, t, reduction="sum")
    loss.backward()
    assert _rel(row.sum(), loss.detach()) < 1e-6
    assert _rel(d, xd.grad / N) < 1e-5
    assert torch.equal(am, x.max(1)[1])
    assert int(am[3]) == 5
    top5 = x.topk(5, 1)[1].eq(t.view(-1, 1)).any(1).int()
    assert torch.equal(t5, top5)
    out = torch.zeros(1, device="cuda")
    ops.reduce_sum(row, N, 1.0 / N, out)
    assert abs(out.item() - loss.item() / N) < 1e-5 * abs(loss.item() / N)


def test_adam_clamp_matches_torch(ops):
    torch.manual_seed(0)
    n = 10000
    p = torch.randn(n, device="cuda")
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=5e-4)
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    import math
    for step in range(1, 6):
        g = torch.randn(n, device="cuda") * (2.0 if step % 2 else 0.1)
        ref.grad = g.clone().clamp_(-0.5, 0.5)
        opt.step()
        gg = g.clone()
        ops.adam_clamp(p, gg, m, v, [(0, 4000), (4000, n - 4000)],
                       [5e-4 / (1 - 0.9 ** step)] * 2, [math.sqrt(1 - 0.999 ** step)] * 2, 0.9, 0.999, 1e-8, 0.5)
        assert torch.equal(gg, ref.grad)      # clamp is applied in place like clip_gradient
    assert _rel(p, ref.data) < 1e-6


@pytest.mark.parametrize("cell", [0, 1])
@pytest.mark.parametrize("B,H,lengths", [
    (5, 16, [7, 6, 4, 3, 3]),
    (96, 512, None),
    (64, 256, "ragged"),
    (130, 32, "ragged"),
])
def test_recurrence_fwd_bwd(ops, cell, B, H, lengths):
    """K3 forward/backward against an unrolled torch (float64) cell on ragged packed batches."""
    import icei_b200
    T = 9
    g = torch.Generator().manual_seed(B * 7 + H + cell)
    if lengths is None:
        lengths = [T] * B
    elif lengths == "ragged":
        lengths = sorted(torch.randint(2, T + 1, (B,), generator=g).tolist(), reverse=True)
        lengths[0] = T
    plan = icei_b200.get_plan(lengths)
    d = plan.dev("cuda")
    N, T = plan.N, plan.T
    XP = (torch.randn(N, 4 * H, generator=g) * 0.7).cuda()
    W = (torch.randn(4 * H, H, generator=g) / H ** 0.5).cuda()
    bhh = (torch.randn(4 * H, generator=g) * 0.1).cuda()
    dH = torch.randn(N, H, generator=g).cuda()
    Hall = torch.empty(N, H, device="cuda"); Call = torch.empty(N, H, device="cuda")
    Hprev = torch.empty(N, H, device="cuda"); gates = torch.empty(N, 4 * H, device="cuda")
    cst = torch.zeros(B, H, device="cuda")
    ops.recur_fwd(cell, H, B, d["bs"], d["off"], 0, T, XP, W, bhh, None, Hall, Call, Hprev, gates, cst)
    dZ = torch.empty(N, 4 * H, device="cuda")
    dh = torch.zeros(B, H, device="cuda"); dc = torch.zeros(B, H, device="cuda")
    ops.recur_bwd(cell, H, B, d["bs"], d["off"], 0, T, W, None, Call, gates, dH, dZ, dh, dc)
    # torch float64 reference
    xp = XP.double().cpu().requires_grad_(True)
    w = W.double().cpu().requires_grad_(True)
    b = bhh.double().cpu()
    h = torch.zeros(B, H, dtype=torch.float64); c = torch.zeros(B, H, dtype=torch.float64)
    hs, hp = [], []
    for t, bt in enumerate(plan.bs):
        z = xp[plan.off[t]:plan.off[t] + bt] + h[:bt] @ w.t() + b
        hp.append(h[:bt])
        if cell == 0:
            i, f, o, gg = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        else:
            i, f, gg, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        c = torch.sigmoid(f) * c[:bt] + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * (c if cell == 0 else torch.tanh(c))
        hs.append(h)
    hall = torch.cat(hs, 0)
    (hall * dH.double().cpu()).sum().backward()
    assert _rel(Hall.cpu(), hall.detach()) < 2e-6
    assert _rel(Hprev.cpu(), torch.cat(hp, 0).detach()) < 2e-6
    assert _rel(dZ.cpu(), xp.grad) < 5e-6
    gw = torch.zeros_like(W)
    ops.gemm(2, dZ, Hprev, gw, 4 * H, H, N, 4 * H, H, H)
    assert _rel(gw.cpu(), w.grad) < 5e-6


def test_recurrence_segmented_equals_one_launch(ops):
    import icei_b200
    B, H = 24, 64
    lengths = sorted([9, 9, 8, 8, 8, 7, 7, 6, 6, 6, 5, 5, 5, 5, 4, 4, 4, 3, 3, 3, 2, 2, 2, 2], reverse=True)
    plan = icei_b200.get_plan(lengths)
    d = plan.dev("cuda")
    N, T = plan.N, plan.T
    torch.manual_seed(3)
    XP = torch.randn(N, 4 * H, device="cuda"); W = torch.randn(4 * H, H, device="cuda") / 8
    outs = []
    for cuts in ([0, T], [0, 1, 2, 5, T]):
        Hall = torch.zeros(N, H, device="cuda"); cst = torch.zeros(B, H, device="cuda")
        for t0, t1 in zip(cuts[:-1], cuts[1:]):
            h_init = Hall[plan.off[t0 - 1]:] if t0 > 0 else None
            ops.recur_fwd(0, H, B, d["bs"], d["off"], t0, t1, XP, W, None, h_init, Hall, None, None, None, cst)
        outs.append(Hall)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("op", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (96, 136, 72), (1920, 2048, 304), (257, 640, 1920),
                                   (2048, 512, 1920), (1920, 10000, 512)])
def test_gemm_tcgen05_bf16(ops, op, M, N, K):
    """tcgen05/TMEM/TMA GEMM vs float64 matmul of the same bf16-rounded operands (fp32-accumulate error only)."""
    g = torch.Generator(device="cuda").manual_seed(M + N * 3 + K * 7 + op)
    pad = lambda v: (v + 7) // 8 * 8
    if op == 2:
        A = torch.randn(K, pad(M), device="cuda", generator=g).bfloat16()
    else:
        A = torch.randn(M, pad(K), device="cuda", generator=g).bfloat16()
    if op == 0:
        B = torch.randn(N, pad(K), device="cuda", generator=g).bfloat16()
    else:
        B = torch.randn(K, pad(N), device="cuda", generator=g).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    C0 = torch.randn(M, N, device="cuda", generator=g)
    C = C0.clone()
    Cb = torch.zeros(M, pad(N), device="cuda", dtype=torch.bfloat16)
    ops.gemm_bf16(op, A, B, M, N, K, A.stride(0), B.stride(0), C=C, ldc=N, Cb=Cb, ldcb=Cb.stride(0), bias=bias, beta=0.5,
                  impl="tc")
    a = A.double()[:, :K] if op != 2 else A.double()[:, :M].t()
    b = B.double()[:, :K].t() if op == 0 else B.double()[:, :N]
    want = a @ b + bias.double() + 0.5 * C0.double()
    assert _rel(C, want) < 1e-5
    assert _rel(Cb[:, :N].float(), want) < 6e-3


def test_gemm_tcgen05_grouped(ops):
    n, F, H = 200, 512, 512
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(n, 4 * F, device="cuda", generator=g).bfloat16()
    B = (torch.randn(4, H, F, device="cuda", generator=g) / 16).bfloat16()
    bias = torch.randn(4 * H, device="cuda", generator=g)
    C = torch.zeros(n, 4 * H, device="cuda")
    ops.gemm_bf16(0, A, B, n, H, F, 4 * F, F, C=C, ldc=4 * H, bias=bias, batch=4, sA=F, sB=H * F, sC=H, sBias=H, impl="tc")
    want = torch.cat([A[:, i * F:(i + 1) * F].double() @ B[i].double().t() + bias[i * H:(i + 1) * H].double()
                      for i in range(4)], 1)
    assert _rel(C, want) < 1e-5


@pytest.mark.parametrize("cell", [0, 1])
@pytest.mark.parametrize("B,H,lengths", [
    (5, 32, [7, 6, 4, 3, 3]),
    (96, 512, None),
    (64, 256, "ragged"),
    (200, 64, "ragged"),
])
def test_recurrence_bf16_fwd_bwd(ops, cell, B, H, lengths):
    """bf16 tensor-core K3 against a float64 unroll that applies the same bf16 rounding to W_hh and to the
    exchanged h / dZ (so the comparison isolates the kernel, tolerance 2e-3) ."""
    import icei_b200
    T = 9
    g = torch.Generator().manual_seed(B * 7 + H + cell + 100)
    if lengths is None:
        lengths = [T] * B
    elif lengths == "ragged":
        lengths = sorted(torch.randint(2, T + 1, (B,), generator=g).tolist(), reverse=True)
        lengths[0] = T
    plan = icei_b200.get_plan(lengths)
    d = plan.dev("cuda")
    N, T = plan.N, plan.T
    XP = (torch.randn(N, 4 * H, generator=g) * 0.7).cuda()
    W = (torch.randn(4 * H, H, generator=g) / H ** 0.5).cuda()
    Wb = W.bfloat16().contiguous()
    bhh = (torch.randn(4 * H, generator=g) * 0.1).cuda()
    dH = torch.randn(N, H, generator=g).cuda()
    Hall = torch.empty(N, H, device="cuda"); Call = torch.empty(N, H, device="cuda")
    Hb = torch.empty(N, H, device="cuda", dtype=torch.bfloat16); Hpb = torch.empty_like(Hb)
    gates = torch.empty(N, 4 * H, device="cuda")
    cst = torch.zeros(B, H, device="cuda")
    ops.recur_fwd_bf16(cell, H, B, d["bs"], d["off"], 0, T, XP, Wb, bhh, None, Hall, Hb, Hpb, Call, gates, cst)
    dZ = torch.empty(N, 4 * H, device="cuda"); dZb = torch.empty(N, 4 * H, device="cuda", dtype=torch.bfloat16)
    dh = torch.zeros(B, H, device="cuda"); dc = torch.zeros(B, H, device="cuda")
    ops.recur_bwd_bf16(cell, H, B, d["bs"], d["off"], 0, T, Wb, None, Call, gates, dH, dZ, dZb, dh, dc)
    torch.cuda.synchronize()
    # reference with the same roundings: h and W_hh rounded to bf16 where they enter the contraction
    w = Wb.double().cpu()
    b = bhh.double().cpu()
    xp = XP.double().cpu()
    h = torch.zeros(B, H, dtype=torch.float64); c = torch.zeros(B, H, dtype=torch.float64)
    hs = []
    for t, bt in enumerate(plan.bs):
        hq = h[:bt].bfloat16().double()
        z = xp[plan.off[t]:plan.off[t] + bt] + hq @ w.t() + b
        if cell == 0:
            i, f, o, gg = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        else:
            i, f, gg, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        c = torch.sigmoid(f) * c[:bt] + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * (c if cell == 0 else torch.tanh(c))
        hs.append(h)
    hall = torch.cat(hs, 0)
    assert _rel(Hall.cpu(), hall) < 2e-3
    assert _rel(Hb.float().cpu(), hall) < 6e-3
    # backward vs the fp32 kernel on the same saved activations (only the dZ exchange is rounded)
    dZ32 = torch.empty(N, 4 * H, device="cuda")
    dh2 = torch.zeros(B, H, device="cuda"); dc2 = torch.zeros(B, H, device="cuda")
    ops.recur_bwd(cell, H, B, d["bs"], d["off"], 0, T, Wb.float(), None, Call, gates, dH, dZ32, dh2, dc2)
    assert _rel(dZ, dZ32) < 1e-2
    assert _rel(dZb.float(), dZ32) < 1.5e-2
    assert _rel(dh, dh2) < 1e-2 and _rel(dc, dc2) < 1e-2


@pytest.mark.parametrize("op", [0, 1, 2])
@pytest.mark.parametrize("M,N,K,batch", [(1920, 512, 300, 1), (1920, 512, 512, 4), (256, 2048, 512, 1), (777, 264, 1920, 1),
                                         (1920, 10000, 512, 1)])
def test_gemm_fp32_on_tensor_cores_split_limbs(ops, op, M, N, K, batch):
    """SN_PREC_BF16X6: fp32 operands as three bf16 limbs, six limb products in one tcgen05 GEMM -> fp32-grade result
    (vs float64 <= 2e-6, the same bound the FFMA kernel is held to), plain and 4-group layouts of the path."""
    g = torch.Generator(device="cuda").manual_seed(M + 7 * N + 13 * K + op)
    if batch == 1:
        A = torch.randn((M, K) if op != 2 else (K, M), device="cuda", generator=g)
        B = torch.randn((N, K) if op == 0 else (K, N), device="cuda", generator=g)
        bias = torch.randn(N, device="cuda", generator=g)
        C0 = torch.randn(M, N, device="cuda", generator=g)
        a = A.double() if op != 2 else A.double().t()
        b = B.double().t() if op == 0 else B.double()
        want = a @ b + bias.double() + 0.5 * C0.double()
        err = {}
        for impl in ("split", "simt"):
            ops.FP32_GEMM[0] = impl
            C = C0.clone()
            before = ops.LAUNCHES[0]
            ops.gemm(op, A, B, C, M, N, K, A.stride(0), B.stride(0), N, bias=bias, beta=0.5)
            err[impl] = _rel(C, want)
            if impl == "split":
                assert ops.LAUNCHES[0] - before >= 3          # two limb expansions + the tensor-core GEMM
        ops.FP32_GEMM[0] = "split"
        print("fp32 GEMM rel error vs float64: split-limb tensor core %.2e, FFMA %.2e" % (err["split"], err["simt"]))
        assert err["simt"] < 2e-6 and err["split"] < 3e-6, err
        return
    # the 4-gate grouped layouts of the factored chain (decoders._input_projection / _input_projection_bwd)
    n, F, H = M, K, N
    if op == 0:      # A2 [n,4F] column groups x U [4][H,F] row-stacked -> XP [n,4H] column groups
        A = torch.randn(n, 4 * F, device="cuda", generator=g); B = torch.randn(4, H, F, device="cuda", generator=g)
        bias = torch.randn(4 * H, device="cuda", generator=g)
        C = torch.zeros(n, 4 * H, device="cuda")
        ops.gemm(0, A, B, C, n, H, F, 4 * F, F, 4 * H, bias=bias, batch=4, sA=F, sB=H * F, sC=H, sBias=H)
        want = torch.cat([A[:, i * F:(i + 1) * F].double() @ B[i].double().t() + bias[i * H:(i + 1) * H].double() for i in range(4)], 1)
    elif op == 1:    # dZ [n,4H] column groups x U [4][H,F] row-stacked (K = H rows) -> dA2 [n,4F]
        A = torch.randn(n, 4 * H, device="cuda", generator=g); B = torch.randn(4, H, F, device="cuda", generator=g)
        C = torch.zeros(n, 4 * F, device="cuda")
        ops.gemm(1, A, B, C, n, F, H, 4 * H, F, 4 * F, batch=4, sA=H, sB=H * F, sC=F)
        want = torch.cat([A[:, i * H:(i + 1) * H].double() @ B[i].double() for i in range(4)], 1)
    else:            # dZ [n,4H]^T x A2 [n,4F] (both column groups, K = n rows) -> gU [4][H,F] row-stacked
        A = torch.randn(n, 4 * H, device="cuda", generator=g); B = torch.randn(n, 4 * F, device="cuda", generator=g)
        C = torch.zeros(4, H, F, device="cuda")
        ops.gemm(2, A, B, C, H, F, n, 4 * H, 4 * F, F, batch=4, sA=H, sB=F, sC=H * F)
        want = torch.stack([A[:, i * H:(i + 1) * H].double().t() @ B[:, i * F:(i + 1) * F].double() for i in range(4)], 0)
    assert _rel(C, want) < 3e-6
