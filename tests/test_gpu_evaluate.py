"""Batched evaluation on the GPU against the reference's own definitions computed with torch
from the full logits: token loss, top-5 accuracy (stylenet/utils.py:127-140), arg-max hypotheses
(train_multitask.py:304-326) and the per-image sample() loop (evaluator.py:63-101)."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


def _accuracy(scores, targets, k):                      # stylenet/utils.py:127-140, verbatim semantics
    batch_size = targets.size(0)
    _, ind = scores.topk(k, 1, True, True)
    correct = ind.eq(targets.view(-1, 1).expand_as(ind))
    return correct.view(-1).float().sum().item() * (100.0 / batch_size)


@pytest.mark.parametrize("kind", ["factored", "nic", "factored_att"])
def test_validate_matches_reference_definitions(kind):
    import icei_b200 as sn
    from icei_b200.evaluate import validate
    from oracle import port
    V, E, H, F, A, D = 331, 28, 32, 40, 24, 48
    torch.manual_seed(0)
    if kind == "factored":
        dec, kw, att = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0), {"mode": "happy"}, False
    elif kind == "nic":
        dec, kw, att = sn.DecoderRNN(E, H, V, 1, dropout=0.0), {}, False
    else:
        dec, kw, att = sn.DecoderFactoredLSTMAtt(A, E, H, F, V, 1, feature_size=D, dropout=0.0), {"mode": "sad"}, True
    port.sharpen_for_decode(dec, scale=20.0)
    dec = dec.cuda().eval()
    batches, want_loss, want_hit, want_tok, want_hyp = [], 0.0, 0.0, 0, []
    for seed, B in ((1, 7), (2, 12)):
        cap, lens, feats = port.synthetic_batch(B, 9, V, E=None if att else E, feat_shape=(3, 3, D) if att else None,
                                                ragged=True, seed=seed)
        all_caps = [[c[:l].tolist(), c[:l].tolist()[::-1]] for c, l in zip(cap, lens)]
        batches.append((feats.cuda(), cap.cuda(), lens, all_caps))
        with torch.no_grad():
            random.seed(0)
            if att:
                l1 = [l - 1 for l in lens]
                out, alphas = dec(cap[:, :-1].cuda(), l1, feats.cuda(), teacher_forcing_ratio=0.0, **kw)
                tgt = port.pack_targets(cap[:, 1:], l1).cuda()
                plens = l1
            else:
                out = dec(cap.cuda(), lens, feats.cuda(), teacher_forcing_ratio=0.0, **kw)
                tgt = port.pack_targets(cap, lens).cuda()
                plens = lens
        n = sum(plens)
        loss = torch.nn.functional.cross_entropy(out, tgt)
        if att:                                          # train_multitask_att.py:321-323 (validation keeps the regulariser)
            loss = loss + ((1.0 - alphas.sum(dim=1)) ** 2).mean()
        want_loss += loss.item() * n
        want_hit += _accuracy(out, tgt, 5) * n
        want_tok += n
        packed = torch.nn.utils.rnn.PackedSequence(out.argmax(1).cpu(), torch.tensor(port.batch_sizes_of(plens)))
        ids, ll = torch.nn.utils.rnn.pad_packed_sequence(packed, batch_first=True)
        for row, l in zip(ids, ll):
            want_hyp.append([w for w in row[:l].tolist() if w not in (1, 2)])
    res = validate(dec, batches, 1, 2, **kw)
    assert res["n_tokens"] == want_tok
    assert abs(res["loss"] - want_loss / want_tok) < 1e-5 * abs(want_loss / want_tok)
    assert abs(res["top5"] - want_hit / want_tok) < 1e-4
    assert res["hypotheses"] == want_hyp
    assert len(res["references"]) == 19 and all(1 not in r and 2 not in r for refs in res["references"] for r in refs)


def test_generate_equals_per_image_sample_loop():
    import icei_b200 as sn
    from icei_b200.evaluate import generate
    from oracle import port
    V, E, H, F = 331, 28, 32, 40
    torch.manual_seed(3)
    dec = sn.DecoderFactoredLSTM(E, H, F, V, 1, dropout=0.0, max_seq_length=12)
    port.lively_weights(dec, end_bias=3.0)
    dec = dec.cuda().eval()
    feats = [torch.randn(5, E).cuda(), torch.randn(3, E).cuda()]
    got = generate(dec, feats, 1, 2, k=3, mode="happy", feed_image=True)
    want = [dec.sample(f[i:i + 1], 1, 2, k=3, mode="happy", feed_image=True)[0].tolist() for f in feats for i in range(f.shape[0])]
    assert got == want and len({len(w) for w in want}) > 1
