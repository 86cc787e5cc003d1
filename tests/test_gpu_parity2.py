"""Module-level parity (GPU), second part: the cases the round-1 review found untested.
  * forward_step of all four decoders against the oracle's forward_step
  * configs[2] (attention over a 7x7x2048 map, B=96, V=10000) at FULL size, fp32 (1e-5) and bf16 (2e-2)
  * batched beam search (n_img > 1) at V=10000 / H=512 against per-image oracle sample(), ids bit-exact
  * bf16-mode greedy ids: agreement with the float64 oracle (a figure, with a floor)
  * whole-object checkpointing after the decode caches exist; learning-rate changes under graph replay
Tolerances are the north star's: fp32 1e-5, bf16 2e-2, ids bit-exact."""
import io
import random

import pytest
import torch

from golden_util import rel_l2

pytestmark = pytest.mark.gpu


def _pair(which, dims, dtype=torch.float64, seed=0, max_seq_length=40):
    """(oracle, drop-in) with identical reference-initialised weights; oracle in ``dtype``."""
    import icei_b200 as sn
    from oracle import port
    A, E, H, F, V, D = dims
    torch.manual_seed(seed)
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        if which == "factored":
            ref = port.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0, max_seq_length=max_seq_length)
        elif which == "nic":
            ref = port.DecoderRNN(E, H, V, 1, dropout=0.0, max_seq_length=max_seq_length)
        elif which == "factored_att":
            ref = port.DecoderFactoredLSTMAtt(A, E, H, F, V, 1, feature_size=D, dropout=0.0, max_seq_length=max_seq_length)
        else:
            ref = port.DecoderRNNAtt(A, E, H, V, 1, feature_size=D, dropout=0.0, max_seq_length=max_seq_length)
    finally:
        torch.set_default_dtype(old)
    if which == "factored":
        dec = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0, max_seq_length=max_seq_length)
    elif which == "nic":
        dec = sn.DecoderRNN(E, H, V, 1, dropout=0.0, max_seq_length=max_seq_length)
    elif which == "factored_att":
        dec = sn.DecoderFactoredLSTMAtt(A, E, H, F, V, 1, feature_size=D, dropout=0.0, max_seq_length=max_seq_length)
    else:
        dec = sn.DecoderRNNAtt(A, E, H, V, 1, feature_size=D, dropout=0.0, max_seq_length=max_seq_length)
    return ref, dec


class _f64:
    """The oracle creates its zero states with the default dtype (like the reference, model.py:176-177)."""

    def __enter__(self):
        self.old = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)

    def __exit__(self, *a):
        torch.set_default_dtype(self.old)


def _sync_weights(ref, dec):
    dec.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    return dec.cuda()


# ---- a3: forward_step ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which,dims,R", [
    ("factored", (0, 300, 512, 512, 1000, 0), 96),
    ("factored", (0, 28, 32, 48, 211, 0), 5),
    ("nic", (0, 300, 512, 0, 1000, 0), 64),
    ("factored_att", (64, 44, 64, 72, 300, 128), 7),
    ("nic_att", (64, 44, 64, 0, 300, 128), 7),
])
def test_forward_step_matches_oracle(which, dims, R):
    """forward_step(embedded, states[, mode]) (stylenet/model.py:115-155, nic/model.py:74-79) on random rows."""
    ref, dec = _pair(which, dims)
    dec = _sync_weights(ref, dec).eval()
    A, E, H, F, V, D = dims
    Ein = E + (D if which.endswith("att") else 0)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(R, Ein, generator=g, dtype=torch.float64)
    h = torch.randn(R, H, generator=g, dtype=torch.float64) * 0.5
    c = torch.randn(R, H, generator=g, dtype=torch.float64) * 0.5
    modes = ["factual", "sad"] if which.startswith("factored") else [None]
    for mode in modes:
        args = (mode,) if mode is not None else ()
        with torch.no_grad():
            hr, (hr2, cr) = ref.forward_step(x, (h, c), *args)
            hg, (hg2, cg) = dec.forward_step(x.float().cuda(), (h.float().cuda(), c.float().cuda()), *args)
        assert hg.shape == (R, H) and hg2.data_ptr() == hg.data_ptr()
        assert rel_l2(hg.cpu(), hr) < 1e-5 and rel_l2(cg.cpu(), cr) < 1e-5
    # chaining two steps feeds the returned states back in (what sample() does)
    with torch.no_grad():
        _, st_r = ref.forward_step(x, (h, c), *args)
        _, st_r = ref.forward_step(x.flip(0), st_r, *args)
        _, st_g = dec.forward_step(x.float().cuda(), (h.float().cuda(), c.float().cuda()), *args)
        _, st_g = dec.forward_step(x.float().flip(0).cuda(), st_g, *args)
    assert rel_l2(st_g[0].cpu(), st_r[0]) < 1e-5 and rel_l2(st_g[1].cpu(), st_r[1]) < 1e-5
    if which == "factored":
        with pytest.raises(ValueError):
            dec.forward_step(x.float().cuda(), (h.float().cuda(), c.float().cuda()), "cheerful")


# ---- a9: configs[2] at full size ------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_attention_config2_full_size(precision, tol):
    """BASELINE.json configs[2]: DecoderFactoredLSTMAtt(512, 300, 512, 512, 10000) over a 7x7x2048 map, B=96, T=20
    (ragged), against the float64 oracle: loss, logits, alphas and EVERY gradient within the north-star tolerance."""
    from oracle import port
    dims = (512, 300, 512, 512, 10000, 2048)
    ref, dec = _pair("factored_att", dims, seed=2)
    dec = _sync_weights(ref, dec).train().set_precision(precision)
    B, T, S = 96, 20, 7
    cap, lens, feats = port.synthetic_batch(B, T, dims[4], feat_shape=(S, S, dims[5]), ragged=True, seed=6)
    l1 = [l - 1 for l in lens]
    tgt = port.pack_targets(cap[:, 1:], l1)
    out_ref, al_ref = ref(cap[:, :-1], l1, feats.double(), teacher_forcing_ratio=1.0, mode="happy")
    loss_ref = port.caption_loss(out_ref, tgt, al_ref)
    ref.zero_grad(); loss_ref.backward()
    out, al = dec(cap[:, :-1].cuda(), l1, feats.cuda(), teacher_forcing_ratio=1.0, mode="happy")
    loss = port.caption_loss(out, tgt.cuda(), al)
    dec.zero_grad(); loss.backward()
    assert rel_l2(out.detach().cpu(), out_ref.detach()) < tol
    assert rel_l2(al.detach().cpu(), al_ref.detach()) < tol
    assert abs(loss.item() - loss_ref.item()) < tol * abs(loss_ref.item())
    gref = {n: p.grad for n, p in ref.named_parameters()}
    # fp32 noise floor: at this size the gradients of the reference's OWN fp32 arithmetic (the same module on torch CPU
    # fp32) sit 3e-5 (median) to 1e-4 (attention net) away from the float64 result -- 20 steps of cancelling sums.  No
    # fp32 implementation can be closer to float64 than that, so in fp32 mode a gradient passes when it is within
    # 1e-5 OR not further from float64 than twice the reference-fp32 module is (measured here, same inputs).
    floor = {}
    if precision == "fp32":
        torch.manual_seed(2)
        ref32 = port.DecoderFactoredLSTMAtt(*dims[:5], 1, feature_size=dims[5], dropout=0.0)
        ref32.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
        o32, a32 = ref32(cap[:, :-1], l1, feats, teacher_forcing_ratio=1.0, mode="happy")
        port.caption_loss(o32, tgt, a32).backward()
        floor = {n: rel_l2(p.grad, gref[n]) for n, p in ref32.named_parameters() if p.grad is not None}
    bad = {}
    for n, p in dec.named_parameters():
        if gref[n] is None:
            assert p.grad is None, n
        elif n.endswith("full_att.bias"):
            assert float(p.grad.abs().max()) < 1e-6          # true gradient is 0 (softmax shift invariance)
        else:
            e = rel_l2(p.grad.cpu(), gref[n])
            if e >= max(tol, 2.0 * floor.get(n, 0.0)):
                bad[n] = (e, floor.get(n))
    assert not bad, bad
    # the fused training entry point gives the same loss / gradients
    g_auto = {n: p.grad.clone() for n, p in dec.named_parameters() if p.grad is not None}
    dec.zero_grad()
    loss2, _ = dec.forward_loss(cap[:, :-1].cuda(), l1, feats.cuda(), full_captions=cap.cuda(), mode="happy")
    assert abs(loss2.item() - loss_ref.item()) < tol * abs(loss_ref.item())
    for n, p in dec.named_parameters():
        if p.grad is not None and not n.endswith("full_att.bias"):
            assert rel_l2(p.grad.cpu(), g_auto[n].cpu()) < (1e-5 if precision == "fp32" else tol), n


# ---- a5 / f2: batched beam search at full vocabulary ----------------------------------------------------------------
@pytest.mark.parametrize("which", ["factored", "nic"])
def test_batched_beam_matches_per_image_oracle_full_vocab(which):
    """sample_batch(n_img in {1, 7, 64}) == per-image oracle sample() (float64), ids bit-exact, at V=10000 / H=512
    with the SURVEY section 8c sharpened weights, k = 1 and 5, both variants (image fed / not fed)."""
    from oracle import port
    dims = (0, 300, 512, 512 if which == "factored" else 0, 10000, 0)
    ref, dec = _pair(which, dims, seed=4, max_seq_length=20)
    port.sharpen_for_decode(ref)
    dec = _sync_weights(ref, dec).eval()
    g = torch.Generator().manual_seed(9)
    feats = torch.randn(64, 300, generator=g, dtype=torch.float64)
    kw = {"mode": "happy"} if which == "factored" else {}
    lens_seen = set()
    for feed in (True, False):
        for k in (1, 5):
            n_ref = 64 if feed else 7          # without the image every caption is the same: 7 images are plenty
            want = []
            with torch.no_grad(), _f64():
                for i in range(n_ref):
                    want.append(ref.sample(feats[i:i + 1], 1, 2, k=k, feed_image=feed, **kw))
            lens_seen |= {w.shape[1] for w in want}
            for n_img in (1, 7, 64):
                if n_img > n_ref:
                    continue
                got = dec.sample_batch(feats[:n_img].float().cuda(), 1, 2, k=k, feed_image=feed, **kw)
                assert len(got) == n_img
                for i in range(n_img):
                    assert got[i].dtype == torch.int64 and got[i].shape[0] == 1
                    assert torch.equal(got[i].cpu(), want[i]), (feed, k, n_img, i, got[i].tolist(), want[i].tolist())
            one = dec.sample(feats[3:4].float().cuda(), 1, 2, k=k, feed_image=feed, **kw)
            assert torch.equal(one.cpu(), want[3])
    assert len(lens_seen) >= 3, lens_seen           # the recipe must produce captions of varied length


@pytest.mark.parametrize("which", ["factored_att", "nic_att"])
def test_batched_attention_beam_matches_per_image_oracle(which):
    """Attention decoders: sample_batch over 6 images == per-image oracle sample() (float64) at configs[2] widths
    (A=512, H=512, D=2048, 7x7 map; vocabulary 2000), k = 1, 3, 5."""
    from oracle import port
    dims = (512, 300, 512, 512 if which == "factored_att" else 0, 2000, 2048)
    ref, dec = _pair(which, dims, seed=5, max_seq_length=16)
    port.sharpen_for_decode(ref)
    dec = _sync_weights(ref, dec).eval()
    g = torch.Generator().manual_seed(10)
    feats = torch.randn(6, 7, 7, 2048, generator=g, dtype=torch.float64)
    kw = {"mode": "sad"} if which == "factored_att" else {}
    for k in (1, 3, 5):
        with torch.no_grad(), _f64():
            want = [ref.sample(feats[i:i + 1], 1, 2, k=k, **kw) for i in range(6)]
        got = dec.sample_batch(feats.float().cuda(), 1, 2, k=k, **kw)
        for i in range(6):
            assert torch.equal(got[i].cpu(), want[i]), (k, i, got[i].tolist(), want[i].tolist())
        one = dec.sample(feats[2:3].float().cuda(), 1, 2, k=k, **kw)
        assert torch.equal(one.cpu(), want[2])


# ---- bf16 greedy agreement ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["factored", "nic"])
def test_bf16_greedy_ids_agree_with_oracle(which):
    """Greedy decode (forward(teacher_forcing_ratio=0), the validation path train_multitask.py:296-299) in bf16 mode
    against the float64 oracle at BASELINE size with the sharpened weights: the fp32 mode must be bit-exact, the
    bf16 mode's agreement is reported and must stay above 0.97 (its logits differ by ~1e-2 relative)."""
    from oracle import port
    dims = (0, 300, 512, 512 if which == "factored" else 0, 10000, 0)
    ref, dec = _pair(which, dims, seed=7)
    port.sharpen_for_decode(ref)
    dec = _sync_weights(ref, dec).eval()
    B, T = 96 if which == "factored" else 64, 20
    cap, lens, feats = port.synthetic_batch(B, T, 10000, E=300, ragged=True, seed=8)
    kw = {"mode": "happy"} if which == "factored" else {}
    with torch.no_grad():
        random.seed(0)
        with _f64():
            ids_ref = ref(cap, lens, feats.double(), teacher_forcing_ratio=0.0, **kw).argmax(1)
        random.seed(0)
        ids32 = dec(cap.cuda(), lens, feats.cuda(), teacher_forcing_ratio=0.0, **kw).argmax(1).cpu()
        dec.set_precision("bf16")
        random.seed(0)
        ids16 = dec(cap.cuda(), lens, feats.cuda(), teacher_forcing_ratio=0.0, **kw).argmax(1).cpu()
    assert torch.equal(ids32, ids_ref)
    agree = (ids16 == ids_ref).float().mean().item()
    print("bf16 greedy agreement vs float64 oracle (%s): %.4f" % (which, agree))
    assert agree > 0.97, agree


# ---- checkpointing with live caches; learning rate under graph replay -------------------------------------------------
def test_pickle_decoder_and_optimizer_after_decode_and_training():
    """torch.save of the whole decoder + optimizers (stylenet/utils.py:62-90) after a graphed greedy forward, a
    sample() and training steps; the reloaded pair continues training exactly like the original."""
    import icei_b200 as sn
    from oracle import port
    V, E, H, F, B, T = 517, 40, 64, 64, 12, 8
    torch.manual_seed(0)
    dec = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0, max_seq_length=10).cuda().train()
    opt = sn.FusedClampAdam(dec, lr=5e-4)
    cap, lens, feats = port.synthetic_batch(B, T, V, E=E, ragged=True, seed=3)
    cap, feats = cap.cuda(), feats.cuda()
    tr = sn.DataParallelTrainer(dec, opt)
    for _ in range(2):
        tr.step(cap, lens, feats, mode="happy")
    with torch.no_grad():
        for _ in range(3):                      # the 2nd identical call captures a CUDA graph
            dec(cap, lens, feats, teacher_forcing_ratio=0.0, mode="happy")
    for _ in range(3):
        dec.sample(feats[:1], 1, 2, k=3, mode="happy", feed_image=True)
    assert dec.__dict__.get("_greedy_graphs") and dec.__dict__.get("_decode_sessions")
    buf = io.BytesIO()
    torch.save({"decoder": dec, "optimizer": opt}, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    dec2, opt2 = back["decoder"], back["optimizer"]
    tr2 = sn.DataParallelTrainer(dec2, opt2)
    for _ in range(2):
        l1, _ = tr.step(cap, lens, feats, mode="sad")
        l2, _ = tr2.step(cap, lens, feats, mode="sad")
    torch.cuda.synchronize()
    assert abs(l1.item() - l2.item()) < 1e-6 * abs(l1.item())
    for (n, p), (_, q) in zip(dec.named_parameters(), dec2.named_parameters()):
        assert rel_l2(q.detach().cpu(), p.detach().cpu()) < 1e-6, n
    assert opt2.step_counts() == opt.step_counts()
    # state_dict route (torch.optim.Adam schema) into a fresh optimizer
    dec3 = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0, max_seq_length=10).cuda().train()
    dec3.load_state_dict(dec.state_dict())
    opt3 = sn.FusedClampAdam(dec3, lr=1.0)
    opt3.load_state_dict(opt.state_dict())
    tr3 = sn.DataParallelTrainer(dec3, opt3)
    l1, _ = tr.step(cap, lens, feats, mode="sad")
    l3, _ = tr3.step(cap, lens, feats, mode="sad")
    for (n, p), (_, q) in zip(dec.named_parameters(), dec3.named_parameters()):
        assert rel_l2(q.detach().cpu(), p.detach().cpu()) < 1e-6, n


def test_graph_replay_follows_learning_rate_changes():
    """adjust_learning_rate (stylenet/utils.py:116-124) mutates param_groups[0]['lr'] between epochs: a replayed
    CUDA graph of the training step must pick the new value up (the lr lives in device memory)."""
    import icei_b200 as sn
    from oracle import port
    V, E, H, F, B, T = 517, 40, 64, 64, 12, 8
    cap, lens, feats = port.synthetic_batch(B, T, V, E=E, ragged=True, seed=3)
    cap, feats = cap.cuda(), feats.cuda()

    def make():
        torch.manual_seed(0)
        d = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0).cuda().train()
        return d, sn.DataParallelTrainer(d, sn.FusedClampAdam(d, lr=5e-4))
    d1, t1 = make()
    d2, t2 = make()
    g = sn.GraphedTrainStep(t2, cap, lens, feats, warmup=3, mode="happy", teacher_forcing_ratio=1.0)
    for _ in range(3):                                   # the graph's warm-up steps are real training steps
        t1.step(cap, lens, feats, mode="happy", teacher_forcing_ratio=1.0)
    for step in range(4):
        if step == 2:
            for t in (t1, t2):
                t.optimizer.param_groups[0]["lr"] *= 0.8          # what adjust_learning_rate does
        l1, _ = t1.step(cap, lens, feats, mode="happy", teacher_forcing_ratio=1.0)
        l2, _ = g()
    torch.cuda.synchronize()
    for (n, p), (_, q) in zip(d1.named_parameters(), d2.named_parameters()):
        assert rel_l2(q.detach().cpu(), p.detach().cpu()) < 1e-5, n


@pytest.mark.gpu
def test_collapsed_chain_follows_weight_updates():
    """Few-row decode collapses U S V once per call and caches it while the weights are unchanged: an optimizer step
    (our kernels), an in-place torch update and a load_state_dict must each refresh it -- sample() after the change ==
    sample() of a fresh decoder holding the same weights."""
    import icei_b200 as sn
    from oracle import port
    V, E, H, F = 300, 28, 64, 64
    torch.manual_seed(0)
    dec = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0)
    port.sharpen_for_decode(dec)
    dec = dec.cuda()
    feats = torch.randn(3, E, generator=torch.Generator().manual_seed(2)).cuda()
    cap, lens, f2 = port.synthetic_batch(8, 7, V, E=E, ragged=True, seed=4)
    opt = sn.FusedClampAdam(dec, lr=5e-2)

    def fresh_ids():
        d2 = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0).cuda().eval()
        d2.load_state_dict(dec.state_dict())
        return [d2.sample(feats[i:i + 1], 1, 2, k=3, mode="happy", feed_image=True) for i in range(3)]

    def ids():
        dec.eval()
        return [dec.sample(feats[i:i + 1], 1, 2, k=3, mode="happy", feed_image=True) for i in range(3)]

    for change in ("none", "adam", "torch", "load"):
        if change == "adam":
            dec.train()
            dec.zero_grad()
            dec.forward_loss(cap.cuda(), lens, f2.cuda(), mode="happy")
            opt.step()
        elif change == "torch":
            with torch.no_grad():
                dec.U_i.weight.mul_(-1.5)
        elif change == "load":
            sd = {k: (v * 0.5 if k.startswith("S_happy") else v) for k, v in dec.state_dict().items()}
            dec.load_state_dict(sd)
        for _ in range(3):                       # (the 3rd call of a session replays its captured graphs)
            got = ids()
        want = fresh_ids()
        for a, b in zip(got, want):
            assert torch.equal(a.cpu(), b.cpu()), change
