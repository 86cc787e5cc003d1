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
