"""Cluster form of the bf16 recurrence (sn_recur_cl.cu) against (1) a float64 unroll with the same bf16 roundings and
(2) the flag-synchronised persistent kernel (sn_recur_bf16.cu) on identical inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import icei_b200
    return icei_b200.ops


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


class _Impl:
    def __init__(self, ops, impl):
        self.ops, self.impl = ops, impl

    def __enter__(self):
        self.old = self.ops.RECUR_IMPL[0]
        self.ops.RECUR_IMPL[0] = self.impl

    def __exit__(self, *a):
        self.ops.RECUR_IMPL[0] = self.old


def _lengths(B, T, kind, g):
    if kind is None:
        return [T] * B
    if kind == "ragged":
        l = sorted(torch.randint(2, T + 1, (B,), generator=g).tolist(), reverse=True)
        l[0] = T
        return l
    return kind


def _inputs(B, H, T, lengths, seed):
    import icei_b200
    g = torch.Generator().manual_seed(seed)
    lengths = _lengths(B, T, lengths, g)
    plan = icei_b200.get_plan(lengths)
    N = plan.N
    XP = (torch.randn(N, 4 * H, generator=g) * 0.7).cuda()
    W = (torch.randn(4 * H, H, generator=g) / H ** 0.5).cuda()
    bhh = (torch.randn(4 * H, generator=g) * 0.1).cuda()
    dH = torch.randn(N, H, generator=g).cuda()
    return plan, XP, W.bfloat16().contiguous(), bhh, dH


def _run(ops, cell, B, H, plan, XP, Wb, bhh, dH, h0=None, c0=None, segments=None):
    d = plan.dev("cuda")
    N, T = plan.N, plan.T
    f32 = dict(device="cuda", dtype=torch.float32)
    b16 = dict(device="cuda", dtype=torch.bfloat16)
    Hall, Call, gates = torch.full((N, H), 7.0, **f32), torch.full((N, H), 7.0, **f32), torch.full((N, 4 * H), 7.0, **f32)
    Hb, Hpb = torch.full((N, H), 7.0, **b16), torch.full((N, H), 7.0, **b16)
    cst = torch.zeros(B, H, **f32) if c0 is None else c0.clone()
    segments = segments or [(0, T)]
    for (t0, t1) in segments:
        h_init = h0 if t0 == 0 else Hall[plan.off[t0 - 1]:]
        ops.recur_fwd_bf16(cell, H, B, d["bs"], d["off"], t0, t1, XP, Wb, bhh, h_init, Hall, Hb, Hpb, Call, gates, cst)
    dZ, dZb = torch.full((N, 4 * H), 7.0, **f32), torch.full((N, 4 * H), 7.0, **b16)
    dh, dc = torch.zeros(B, H, **f32), torch.zeros(B, H, **f32)
    for (t0, t1) in reversed(segments):
        ops.recur_bwd_bf16(cell, H, B, d["bs"], d["off"], t0, t1, Wb, c0, Call, gates, dH, dZ, dZb, dh, dc)
    torch.cuda.synchronize()
    return dict(Hall=Hall, Call=Call, gates=gates, Hb=Hb, Hpb=Hpb, cst=cst, dZ=dZ, dZb=dZb, dh=dh, dc=dc)


def _unroll(cell, B, H, plan, XP, Wb, bhh, h0=None, c0=None):
    w, b, xp = Wb.double().cpu(), bhh.double().cpu(), XP.double().cpu()
    h = torch.zeros(B, H, dtype=torch.float64) if h0 is None else h0.double().cpu()
    c = torch.zeros(B, H, dtype=torch.float64) if c0 is None else c0.double().cpu()
    hs, hp = [], []
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
        hs.append(h)
    return torch.cat(hs, 0), torch.cat(hp, 0)


@pytest.mark.parametrize("cell", [0, 1])
@pytest.mark.parametrize("B,H,lengths", [
    (96, 512, None),
    (64, 512, "ragged"),
    (5, 512, [7, 6, 4, 3, 3]),
    (150, 512, "ragged"),        # 10 sample slices: more clusters than fit at once -> later waves
    (37, 256, "ragged"),
    (96, 128, "ragged"),
    (1, 512, [9]),
])
def test_cluster_recurrence_matches_reference_and_flag_kernel(ops, cell, B, H, lengths):
    T = 9
    plan, XP, Wb, bhh, dH = _inputs(B, H, T, lengths, B * 7 + H + cell + 300)
    if ops.lib().sn_recur_cl_max_clusters(H) < 1:
        pytest.skip("cluster launch of %d CTAs not available on this device" % (H // 32))
    with _Impl(ops, "cluster"):
        cl = _run(ops, cell, B, H, plan, XP, Wb, bhh, dH)
    with _Impl(ops, "flags"):
        fl = _run(ops, cell, B, H, plan, XP, Wb, bhh, dH)
    hall, hprev = _unroll(cell, B, H, plan, XP, Wb, bhh)
    assert _rel(cl["Hall"].cpu(), hall) < 2e-3
    assert _rel(cl["Hb"].float().cpu(), hall) < 6e-3
    assert _rel(cl["Hpb"].float().cpu(), hprev) < 6e-3
    # both kernels round the same operands to bf16 and accumulate in fp32: they agree far inside the bf16 error
    for k in ("Hall", "Call", "gates", "cst"):
        assert _rel(cl[k], fl[k]) < 2e-3, k
    for k in ("dZ", "dh", "dc"):
        assert _rel(cl[k], fl[k]) < 1e-2, k
    assert _rel(cl["dZb"].float(), fl["dZb"].float()) < 1.5e-2
    # and against the fp32 backward kernel on the cluster kernel's own saved activations
    d = plan.dev("cuda")
    dZ32 = torch.empty_like(cl["dZ"])
    dh2 = torch.zeros(B, H, device="cuda"); dc2 = torch.zeros(B, H, device="cuda")
    ops.recur_bwd(cell, H, B, d["bs"], d["off"], 0, plan.T, Wb.float(), None, cl["Call"], cl["gates"], dH, dZ32, dh2, dc2)
    assert _rel(cl["dZ"], dZ32) < 1e-2
    assert _rel(cl["dh"], dh2) < 1e-2 and _rel(cl["dc"], dc2) < 1e-2


def test_cluster_recurrence_segments_and_initial_state(ops):
    """Segmented launches (the non-teacher-forced path) and a non-zero initial state (the attention decoders' init_h /
    init_c) give the same result as one launch; the returned carries are the gradients w.r.t. that initial state."""
    cell, B, H, T = 0, 40, 512, 8
    if ops.lib().sn_recur_cl_max_clusters(H) < 1:
        pytest.skip("cluster launch not available")
    plan, XP, Wb, bhh, dH = _inputs(B, H, T, "ragged", 77)
    g = torch.Generator().manual_seed(5)
    h0 = (torch.randn(B, H, generator=g) * 0.5).cuda()
    c0 = (torch.randn(B, H, generator=g) * 0.5).cuda()
    with _Impl(ops, "cluster"):
        one = _run(ops, cell, B, H, plan, XP, Wb, bhh, dH, h0, c0)
        seg = _run(ops, cell, B, H, plan, XP, Wb, bhh, dH, h0, c0, segments=[(0, 3), (3, 4), (4, T)])
    with _Impl(ops, "flags"):
        fl = _run(ops, cell, B, H, plan, XP, Wb, bhh, dH, h0, c0)
    hall, _ = _unroll(cell, B, H, plan, XP, Wb, bhh, h0, c0)
    assert _rel(one["Hall"].cpu(), hall) < 2e-3
    for k in ("Hall", "Call", "gates", "cst"):
        assert _rel(seg[k], one[k]) < 1e-6, k            # same arithmetic, only the launch boundaries differ
    for k in ("dZ", "dh", "dc"):
        assert _rel(seg[k], one[k]) < 1e-2, k            # (the carried dh is re-rounded to fp32 at a boundary only)
        assert _rel(one[k], fl[k]) < 1e-2, k


def test_cluster_recurrence_is_deterministic(ops):
    cell, B, H, T = 1, 96, 512, 12
    if ops.lib().sn_recur_cl_max_clusters(H) < 1:
        pytest.skip("cluster launch not available")
    plan, XP, Wb, bhh, dH = _inputs(B, H, T, None, 11)
    with _Impl(ops, "cluster"):
        a = _run(ops, cell, B, H, plan, XP, Wb, bhh, dH)
        b = _run(ops, cell, B, H, plan, XP, Wb, bhh, dH)
    for k in a:
        assert torch.equal(a[k], b[k]), k
