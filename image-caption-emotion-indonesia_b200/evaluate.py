"""Batched evaluation: the two loops that follow the decoder in the reference,
  * ``validate``  = val_factual / val_emotion of stylenet/train_multitask.py:272-361 (teacher_forcing_ratio = 0 forward,
    token loss, top-5 accuracy, per-caption arg-max ids for BLEU), and
  * ``generate``  = the per-image ``sample()`` loop of stylenet/evaluator.py:63-101,
with everything id-level done on the device: the loss / top-5 / arg-max come out of the fused softmax kernel instead of
five passes over the [N, V] logits on the host, and beam search runs for a whole batch of images at once
(``sample_batch``).  Ids -> words -> BLEU stays with the caller (nltk, CPU), exactly the reference's split."""
import torch

from .packing import get_plan


def _strip(ids, start, end):
    return [w for w in ids if w != start and w != end]


@torch.no_grad()
def validate(decoder, batches, start_token, end_token, encoder=None, mode=None, attention=None):
    """``batches`` yields (images_or_features, captions [B,T] int64, lengths, all_captions | None) like the reference's
    data loader (sorted by length, data_loader.py:116-145).  Returns dict(loss, top5 (percent), hypotheses,
    references, n_tokens) with the reference's definitions: loss = token-mean cross entropy averaged over batches
    weighted by tokens (AverageMeter, utils.py:93-110), top5 = accuracy(scores, targets, 5) (utils.py:127-140),
    hypotheses = per-caption arg-max ids cut to the caption length with <start>/<end> removed."""
    was_training = decoder.training
    decoder.eval()
    if encoder is not None:
        encoder.eval()
    att = hasattr(decoder, "attention") if attention is None else attention
    kw = {} if mode is None else {"mode": mode}
    tot_loss = tot_hit = 0.0
    tot_tok = 0
    hyps, refs = [], []
    for images, captions, lengths, all_caps in batches:
        feats = encoder(images) if encoder is not None else images
        lengths = [int(l) for l in lengths]
        if att:
            l1 = [l - 1 for l in lengths]
            loss, st = decoder.forward_loss(captions[:, :-1], l1, feats, full_captions=captions,
                                            teacher_forcing_ratio=0.0, backward=False, **kw)
            plan = get_plan(l1)
        else:
            loss, st = decoder.forward_loss(captions, lengths, feats, teacher_forcing_ratio=0.0, backward=False, **kw)
            plan = get_plan(lengths)
        n = plan.N
        tot_loss += float(loss.item()) * n
        tot_hit += float(st["top5hit"].sum().item())
        tot_tok += n
        am = st["argmax"].cpu().tolist()
        for b, L in enumerate(plan.lengths):
            hyps.append(_strip([am[plan.off[t] + b] for t in range(L)], start_token, end_token))
        if all_caps is not None:
            for caps in all_caps:
                refs.append([_strip([int(w) for w in c], start_token, end_token) for c in caps])
    decoder.train(was_training)
    return {"loss": tot_loss / max(tot_tok, 1), "top5": 100.0 * tot_hit / max(tot_tok, 1), "hypotheses": hyps,
            "references": refs, "n_tokens": tot_tok}


@torch.no_grad()
def generate(decoder, feature_batches, start_token, end_token, k=5, mode=None, encoder=None, **sample_kw):
    """Beam-search captions for every image: ``feature_batches`` yields feature tensors ([n, E] / [n, S, S, D]) or, with
    ``encoder``, image batches.  Returns a list of id lists (one per image, as ``sample()`` would return them)."""
    decoder.eval()
    kw = dict(sample_kw)
    if mode is not None:
        kw["mode"] = mode
    out = []
    for x in feature_batches:
        feats = encoder(x) if encoder is not None else x
        for ids in decoder.sample_batch(feats, start_token, end_token, k=k, **kw):
            out.append(ids[0].tolist())
    return out
