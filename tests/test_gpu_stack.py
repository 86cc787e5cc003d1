"""GPU parity of the multi-layer FactoredLSTM stack (BASELINE.json configs[3]) against
 (a) tests/golden/stack3.npz -- UNMODIFIED reference DecoderFactoredLSTM objects composed by oracle/stack.py
     (oracle/make_golden_stack.py), fp32 mode <= 1e-5, greedy arg-max ids bit-exact, and
 (b) the oracle port at configs[3] dimensions (F = 1024, 3 layers) in bf16 mode <= 2e-2."""
import random

import numpy as np
import pytest
import torch

from golden_util import load_golden, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _layer_sds(rec, L, prefix="sd.", dtype=torch.float32):
    out = []
    for l in range(L):
        pre = "%s%d." % (prefix, l)
        out.append({k[len(pre):]: torch.from_numpy(v).to(dtype) if v.dtype.kind == "f" else torch.from_numpy(v)
                    for k, v in rec.items() if k.startswith(pre)})
    return out


def _build(rec, dropout=0.0):
    import icei_b200 as sn
    V, E, H, F, L = (int(rec["meta." + k]) for k in ("V", "E", "H", "F", "L"))
    m = sn.DecoderFactoredLSTMStack(E, H, F, V, L, dropout=dropout, max_seq_length=12)
    m.load_layer_state_dicts(_layer_sds(rec, L))
    return m.cuda(), L


def _inputs(rec):
    cap = torch.from_numpy(rec["in.captions"]).cuda()
    lens = [int(x) for x in rec["in.lengths"]]
    feats = torch.from_numpy(rec["in.features"]).float().cuda()
    return cap, lens, feats


def _golden_name(n):
    """stack parameter name -> (layer, reference name)"""
    if n.startswith("l") and n[1].isdigit() and n[2] == "_":
        return int(n[1]), n[3:]
    return 0, n


def _check_grads(dec, rec, pre):
    want = {k[len(pre):] for k in rec if k.startswith(pre)}
    got = set()
    for n, p in dec.named_parameters():
        if p.grad is None:
            continue
        l, rn = _golden_name(n)
        key = "%d.%s" % (l, rn)
        got.add(key)
        assert rel_l2(p.grad.cpu(), rec[pre + key]) < TOL, n
    assert got == want, "set of parameters with a gradient"


def test_stack_golden_forward_backward():
    from oracle import port
    rec = load_golden("stack3")
    dec, L = _build(rec)
    dec.train()
    cap, lens, feats0 = _inputs(rec)
    tgt = port.pack_targets(cap.cpu(), lens).cuda()
    for mode in ("factual", "sad"):
        dec.zero_grad()
        feats = feats0.clone().requires_grad_(True)
        random.seed(1234)
        out = dec(cap, lens, feats, teacher_forcing_ratio=1.0, mode=mode)
        loss = torch.nn.functional.cross_entropy(out, tgt)
        loss.backward()
        assert rel_l2(out.detach().cpu(), rec["tf1.logits." + mode]) < TOL
        assert abs(loss.item() - float(rec["tf1.loss." + mode])) < TOL * abs(float(rec["tf1.loss." + mode]))
        assert rel_l2(feats.grad.cpu(), rec["tf1.dfeatures." + mode]) < TOL
        _check_grads(dec, rec, "tf1.grad.%s." % mode)
        with torch.no_grad():
            random.seed(1234)
            out0 = dec(cap, lens, feats0, teacher_forcing_ratio=0.0, mode=mode)
        assert rel_l2(out0.cpu(), rec["tf0.logits." + mode]) < TOL
        assert np.array_equal(out0.argmax(1).cpu().numpy(), rec["tf0.argmax." + mode])
        dec.zero_grad()
        random.seed(1234)
        out5 = dec(cap, lens, feats0, teacher_forcing_ratio=0.5, mode=mode)
        torch.nn.functional.cross_entropy(out5, tgt).backward()
        assert rel_l2(out5.detach().cpu(), rec["tf05.logits." + mode]) < TOL
        _check_grads(dec, rec, "tf05.grad.%s." % mode)


def test_stack_multitask_adam_schedule():
    """forward_loss + two FusedClampAdam objects alternating factual / emotion passes over the same stack
    (train_multitask.py:192-235) == reference clip_gradient + torch.optim.Adam after 4 steps."""
    import icei_b200 as sn
    rec = load_golden("stack3")
    dec, L = _build(rec)
    dec.train()
    cap, lens, feats = _inputs(rec)
    opt_a = sn.FusedClampAdam(dec, lr=2e-4, grad_clip=0.5)
    opt_b = sn.FusedClampAdam(dec, lr=5e-4, grad_clip=0.5)
    loss = None
    for opt, mode in [(opt_a, "factual"), (opt_b, "sad"), (opt_a, "factual"), (opt_b, "happy")]:
        dec.zero_grad()
        random.seed(1234)
        loss, _ = dec.forward_loss(cap, lens, feats, teacher_forcing_ratio=1.0, mode=mode)
        opt.step()
    assert abs(loss.item() - float(rec["adam4.loss_last"])) < TOL * abs(float(rec["adam4.loss_last"]))
    for n, p in dec.named_parameters():
        l, rn = _golden_name(n)
        assert rel_l2(p.detach().cpu(), rec["adam4.sd.%d.%s" % (l, rn)]) < TOL, n


def _oracle_stack(E, H, F, V, L, seed):
    from oracle import port
    torch.manual_seed(seed)
    return [port.DecoderFactoredLSTM(E if l == 0 else H, H, F, V, 1, dropout=0.0) for l in range(L)]


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_stack_config4_dimensions_vs_oracle(precision, tol):
    """3 layers, factored 1024, hidden 512, embed 300 (configs[3]) on a reduced batch / vocabulary the CPU oracle
    finishes in seconds; loss and every gradient within the north-star tolerance of the mode."""
    import icei_b200 as sn
    from oracle import port
    from oracle.stack import stack_forward
    E, H, F, V, L, B, T = 300, 512, 1024, 1000, 3, 24, 9
    layers = _oracle_stack(E, H, F, V, L, 3)
    dec = sn.DecoderFactoredLSTMStack(E, H, F, V, L, dropout=0.0)
    dec.load_layer_state_dicts([l.state_dict() for l in layers])
    dec = dec.cuda().set_precision(precision)
    dec.train()
    cap, lens, feats = port.synthetic_batch(B, T, V, E=E, ragged=True, seed=9)
    tgt = port.pack_targets(cap, lens)
    random.seed(1234)
    out = stack_forward(layers, cap, lens, feats, teacher_forcing_ratio=1.0, mode="happy")
    loss = torch.nn.functional.cross_entropy(out, tgt)
    loss.backward()
    random.seed(1234)
    loss_g, stats = dec.forward_loss(cap.cuda(), lens, feats.cuda(), teacher_forcing_ratio=1.0, mode="happy")
    assert abs(loss_g.item() - loss.item()) < tol * abs(loss.item())
    if precision == "fp32":
        assert torch.equal(stats["argmax"].cpu(), out.argmax(1))
    ref = {}
    for l, layer in enumerate(layers):
        for n, p in layer.named_parameters():
            if p.grad is not None:
                ref[("" if l == 0 else "l%d_" % l) + n] = p.grad
    got = {n for n, p in dec.named_parameters() if p.grad is not None}
    assert got == set(ref)
    for n, p in dec.named_parameters():
        if p.grad is not None:
            assert rel_l2(p.grad.cpu(), ref[n]) < tol, n


def test_early_vocab_adam_equals_single_step():
    """The single-GPU trainer updates the vocabulary projection on a side stream as soon as dC is final and the rest at
    the end (optimizer.step(only=...) + step(skip=...)): parameters after 3 steps == one optimizer.step() per step."""
    import icei_b200 as sn
    from oracle import port
    E, H, F, V, B, T = 40, 64, 64, 517, 56, 8
    cap, lens, feats = port.synthetic_batch(B, T, V, E=E, ragged=True, seed=3)
    cap, feats = cap.cuda(), feats.cuda()
    outs = []
    for early in (True, False):
        torch.manual_seed(5)
        dec = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0).cuda().set_precision("bf16")
        dec.train()
        opt = sn.FusedClampAdam(dec, lr=1e-3, grad_clip=0.5)
        tr = sn.DataParallelTrainer(dec, opt)
        for _ in range(3):
            random.seed(1)
            if early:
                loss, _ = tr.step(cap, lens, feats, teacher_forcing_ratio=1.0, mode="sad")
            else:
                for p in dec.parameters():
                    p.grad = None
                loss, _ = dec.forward_loss(cap, lens, feats, teacher_forcing_ratio=1.0, mode="sad")
                opt.step()
        torch.cuda.synchronize()
        outs.append((loss.item(), {n: p.detach().clone() for n, p in dec.named_parameters()}, opt.step_counts()))
    assert abs(outs[0][0] - outs[1][0]) < 1e-5 * abs(outs[1][0])
    assert outs[0][2] == outs[1][2]
    # not bit-exact by construction: the embedding / bias gradients are accumulated with fp32 atomics
    worst = max((rel_l2(outs[0][1][n].cpu(), outs[1][1][n].cpu()), n) for n in outs[0][1])
    assert worst[0] < 1e-5, worst


def test_stack_forward_step_and_beam_sample_vs_oracle_composition():
    """forward_step (states [L, R, H]) and sample() / sample_batch() THROUGH the stack against the oracle composition
    (oracle/stack.py::stack_forward_step / stack_sample over port layers, float64): states <= 1e-5, ids bit-exact."""
    import icei_b200 as sn
    from oracle import port
    from oracle.stack import stack_forward_step, stack_sample
    E, H, F, V, L = 44, 64, 96, 517, 3
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(5)
        layers = [port.DecoderFactoredLSTM(E if l == 0 else H, H, F, V, 1, dropout=0.0, max_seq_length=14) for l in range(L)]
        for l, layer in enumerate(layers):
            port.lively_weights(layer, end_bias=2.0, seed=5 + l)
        torch.set_default_dtype(torch.float32)
        dec = sn.DecoderFactoredLSTMStack(E, H, F, V, L, dropout=0.0, max_seq_length=14)
        torch.set_default_dtype(torch.float64)          # (the oracle creates its zero states with the default dtype)
        dec.load_layer_state_dicts([{k: v.float() for k, v in l.state_dict().items()} for l in layers])
        dec = dec.cuda().eval()
        g = torch.Generator().manual_seed(1)
        for R in (5, 40):                      # matrix-vector kernels / the batched path
            x = torch.randn(R, E, generator=g)
            h = torch.randn(L, R, H, generator=g) * 0.5
            c = torch.randn(L, R, H, generator=g) * 0.5
            with torch.no_grad():
                top_r, (h_r, c_r) = stack_forward_step(layers, x, (h, c), "sad")
                top_g, (h_g, c_g) = dec.forward_step(x.float().cuda(), (h.float().cuda(), c.float().cuda()), "sad")
            assert tuple(h_g.shape) == (L, R, H)
            assert rel_l2(top_g.cpu(), top_r) < 1e-5 and rel_l2(h_g.cpu(), h_r) < 1e-5 and rel_l2(c_g.cpu(), c_r) < 1e-5
        feats = torch.randn(20, E, generator=g)
        lens = set()
        for k in (1, 3, 5):
            with torch.no_grad():
                want = [stack_sample(layers, feats[i:i + 1], 1, 2, k=k, mode="happy", feed_image=True) for i in range(20)]
            lens |= {w.shape[1] for w in want}
            got = dec.sample_batch(feats.float().cuda(), 1, 2, k=k, mode="happy", feed_image=True)     # 20 x k rows
            for i in range(20):
                assert torch.equal(got[i].cpu(), want[i]), (k, i)
            for i in (0, 7):                                                                         # <= 16 rows
                one = dec.sample(feats[i:i + 1].float().cuda(), 1, 2, k=k, mode="happy", feed_image=True)
                assert torch.equal(one.cpu(), want[i]), (k, i)
        assert len(lens) >= 2
    finally:
        torch.set_default_dtype(old)
