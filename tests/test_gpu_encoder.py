"""Encoder tail -> decoder hand-off against the reference modules' own tail
(stylenet/model.py:19-26 Linear + BatchNorm1d(momentum=0.01); stylenet/model_att.py:19-28 AdaptiveAvgPool2d + permute),
built here from the same torch layers the reference uses (its constructor downloads ResNet weights, so the trunk is
replaced by a small stand-in; the tail is what is under test)."""
import pytest
import torch
import torch.nn as nn

from golden_util import rel_l2

pytestmark = pytest.mark.gpu


class _TinyTrunk(nn.Module):
    """Stand-in for ResNet-152 without its last layers: images [B,3,H,W] -> [B, C, H/8, W/8]."""

    def __init__(self, C, pool):
        super().__init__()
        self.conv = nn.Conv2d(3, C, 8, stride=8)
        self.pool = nn.AdaptiveAvgPool2d((1, 1)) if pool else nn.Identity()

    def forward(self, x):
        return self.pool(torch.relu(self.conv(x)))


@pytest.mark.parametrize("B,C,E", [(96, 2048, 300), (5, 64, 12)])
def test_encoder_cnn_tail_matches_linear_batchnorm(B, C, E):
    import icei_b200 as sn
    torch.manual_seed(0)
    trunk = _TinyTrunk(C, True).cuda()
    enc = sn.EncoderCNN(E, backbone=trunk, in_features=C).cuda()
    assert [n for n, _ in enc.named_parameters() if not n.startswith("resnet.")] == ["linear.weight", "linear.bias", "bn.weight", "bn.bias"]
    ref_lin = nn.Linear(C, E).cuda().double()
    ref_bn = nn.BatchNorm1d(E, momentum=0.01).cuda().double()
    with torch.no_grad():
        enc.bn.weight.uniform_(0.5, 1.5); enc.bn.bias.uniform_(-0.2, 0.2)
    ref_lin.load_state_dict({k: v.double() for k, v in enc.linear.state_dict().items()})
    ref_bn.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in enc.bn.state_dict().items()})
    for it in range(3):                                   # running statistics accumulate across steps
        img = torch.randn(B, 3, 32, 32, device="cuda")
        with torch.no_grad():
            f = trunk(img).reshape(B, -1).double()
        want = ref_bn(ref_lin(f))
        got = enc(img)
        assert got.shape == (B, E) and rel_l2(got.detach().cpu(), want.detach().cpu()) < 1e-5
        w = torch.randn(B, E, device="cuda")
        ref_lin.zero_grad(); ref_bn.zero_grad(); enc.zero_grad()
        (want * w.double()).sum().backward()
        (got * w).sum().backward()
        for a, b in ((enc.linear.weight, ref_lin.weight), (enc.linear.bias, ref_lin.bias), (enc.bn.weight, ref_bn.weight),
                     (enc.bn.bias, ref_bn.bias)):
            # (d linear.bias is analytically zero behind a batch norm: compare absolutely there)
            if a is enc.linear.bias:
                assert float(a.grad.abs().max()) < 1e-4 * float(enc.linear.weight.grad.abs().max() + 1e-30)
            else:
                assert rel_l2(a.grad.cpu(), b.grad.cpu()) < 1e-5
        assert all(p.grad is None for p in trunk.parameters())          # trunk runs under no_grad like the reference
    assert rel_l2(enc.bn.running_mean.cpu(), ref_bn.running_mean.cpu()) < 1e-5
    assert rel_l2(enc.bn.running_var.cpu(), ref_bn.running_var.cpu()) < 1e-5
    assert int(enc.bn.num_batches_tracked) == 3
    enc.eval(); ref_bn.eval()
    img = torch.randn(7, 3, 32, 32, device="cuda")
    with torch.no_grad():
        want = ref_bn(ref_lin(trunk(img).reshape(7, -1).double()))
        got = enc(img)
    assert rel_l2(got.cpu(), want.cpu()) < 1e-5


@pytest.mark.parametrize("B,C,hw,S", [(96, 2048, 7, 7), (4, 2048, 7, 14), (3, 100, 14, 7), (2, 64, 9, 4), (2, 40, 5, 3)])
def test_attention_encoder_tail_matches_pool_permute(B, C, hw, S):
    import icei_b200 as sn
    torch.manual_seed(1)
    trunk = _TinyTrunk(C, False).cuda()
    enc = sn.EncoderCNNAtt(S, backbone=trunk, emit_bf16=True).cuda()
    img = torch.randn(B, 3, hw * 8, hw * 8, device="cuda")
    with torch.no_grad():
        fmap = trunk(img)
    want = nn.AdaptiveAvgPool2d((S, S))(fmap.double()).permute(0, 2, 3, 1)      # the reference's tail, verbatim
    got = enc(img)
    assert got.shape == (B, S, S, C) and got.is_contiguous()
    assert rel_l2(got.cpu(), want.cpu()) < 1e-6
    assert rel_l2(got.sn_mean.cpu(), want.reshape(B, -1, C).mean(1).cpu()) < 1e-6
    assert rel_l2(got.sn_bf16.float().cpu(), want.cpu()) < 4e-3
    # backward of the tail alone (the trunk is frozen in the reference; fine-tuning would use this)
    x = fmap.clone().requires_grad_(True)
    xd = fmap.double().clone().requires_grad_(True)
    from icei_b200.encoders import _PoolNHWCFn
    w = torch.randn(B, S, S, C, device="cuda")
    (_PoolNHWCFn.apply(x, S, False)[0] * w).sum().backward()
    (nn.AdaptiveAvgPool2d((S, S))(xd).permute(0, 2, 3, 1) * w.double()).sum().backward()
    assert rel_l2(x.grad.cpu(), xd.grad.cpu()) < 1e-6


def test_attention_decoder_consumes_the_fused_tail():
    """features from EncoderCNNAtt (contiguous NHWC + attached mean / bf16 copy) give the decoder the same result as
    the reference-style permuted view of the pooled NCHW map."""
    import icei_b200 as sn
    from oracle import port
    torch.manual_seed(2)
    A, E, H, F, V, D, S, B, T = 64, 40, 64, 64, 300, 128, 5, 9, 7
    trunk = _TinyTrunk(D, False).cuda()
    enc = sn.EncoderCNNAtt(S, backbone=trunk, emit_bf16=True).cuda()
    cap, lens, _ = port.synthetic_batch(B, T, V, ragged=True, seed=4)
    l1 = [l - 1 for l in lens]
    img = torch.randn(B, 3, 40, 40, device="cuda")
    with torch.no_grad():
        ref_feats = nn.AdaptiveAvgPool2d((S, S))(trunk(img)).permute(0, 2, 3, 1)       # non-contiguous view
    for precision, tol in (("fp32", 1e-6), ("bf16", 1e-6)):
        dec = sn.DecoderFactoredLSTMAtt(A, E, H, F, V, 1, feature_size=D, dropout=0.0).cuda().eval().set_precision(precision)
        with torch.no_grad():
            o1, a1 = dec(cap[:, :-1].cuda(), l1, enc(img), teacher_forcing_ratio=1.0, mode="sad")
            o2, a2 = dec(cap[:, :-1].cuda(), l1, ref_feats, teacher_forcing_ratio=1.0, mode="sad")
        assert rel_l2(o1.cpu(), o2.cpu()) < tol and rel_l2(a1.cpu(), a2.cpu()) < tol
