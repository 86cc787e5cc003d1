"""Checkpoint interchange with the reference.

The reference checkpoints are WHOLE-OBJECT pickles (stylenet/utils.py:63-90: ``{'epoch', 'epochs_since_improvement',
'bleu-4', 'encoder', 'decoder', 'optimizer', 'lang_optimizer'}`` saved with torch.save) which the trainers and the demo
backend load back as live modules (train_multitask.py:169-176, app/backend/model.py:490-499), next to plain
``state_dict`` files (``decoder-8.ckpt``, exporter.py:30-33).  This module

  * writes checkpoints with the reference's function and file naming (``save_checkpoint``),
  * turns a reference checkpoint / module / state_dict into the drop-in modules (``convert_checkpoint``,
    ``decoder_from_reference``) -- constructor arguments are recovered from the parameter shapes, so nothing but the
    tensors is needed -- and torch.optim.Adam objects into ``FusedClampAdam`` with their moments and step counts,
  * and exports drop-in state back into what the reference classes load (``export_state``).
Host logic only (no kernels): it runs wherever torch runs."""
import os

import torch

from .optim import FusedClampAdam


def _kind(sd):
    keys = set(sd)
    if "B.weight" in keys and "attention.encoder_att.weight" in keys:
        return "factored_att"
    if "B.weight" in keys and any(k.startswith("l1_") for k in keys):
        return "stack"
    if "B.weight" in keys:
        return "factored"
    if "embed.weight" in keys and "attention.encoder_att.weight" in keys:
        return "nic_att"
    if "embed.weight" in keys and "lstm.weight_hh" in keys:
        return "nic"
    raise ValueError("not a decoder state_dict of the reference (stylenet/model.py, model_att.py, nic/model.py, "
                     "nic/model_att.py): keys %s..." % sorted(keys)[:5])


def decoder_from_state_dict(sd, dropout=0.22, max_seq_length=40, device=None):
    """Build the matching drop-in decoder for a reference ``state_dict`` and load it."""
    from . import DecoderFactoredLSTM, DecoderFactoredLSTMAtt, DecoderRNN, DecoderRNNAtt, DecoderFactoredLSTMStack
    kind = _kind(sd)
    if kind in ("factored", "factored_att", "stack"):
        V, E = sd["B.weight"].shape
        H, F = sd["U_i.weight"].shape
        if kind == "factored":
            dec = DecoderFactoredLSTM(E, H, F, V, 1, dropout=dropout, max_seq_length=max_seq_length)
        elif kind == "stack":
            layers = 1 + max(int(k[1:k.index("_")]) for k in sd if k.startswith("l") and k[1].isdigit())
            dec = DecoderFactoredLSTMStack(E, H, F, V, layers, dropout=dropout, max_seq_length=max_seq_length)
        else:
            A, D = sd["attention.encoder_att.weight"].shape
            dec = DecoderFactoredLSTMAtt(A, E, H, F, V, 1, feature_size=D, dropout=dropout, max_seq_length=max_seq_length)
    else:
        V, E = sd["embed.weight"].shape
        H = sd["lstm.weight_hh"].shape[1]
        if kind == "nic":
            dec = DecoderRNN(E, H, V, 1, dropout=dropout, max_seq_length=max_seq_length)
        else:
            A, D = sd["attention.encoder_att.weight"].shape
            dec = DecoderRNNAtt(A, E, H, V, 1, feature_size=D, dropout=dropout, max_seq_length=max_seq_length)
    dec.load_state_dict({k: v.detach().float() for k, v in sd.items()})
    return dec.to(device) if device is not None else dec


def decoder_from_reference(module, device=None):
    """Drop-in decoder for a live reference decoder object (e.g. ``checkpoint['decoder']`` un-pickled with the
    reference's classes importable): same weights, dropout probability, max_seq_length and train / eval mode."""
    drop = getattr(getattr(module, "dropout", None), "p", 0.22)
    dec = decoder_from_state_dict(module.state_dict(), dropout=drop,
                                  max_seq_length=getattr(module, "max_seq_length", 40), device=device)
    dec.train(module.training)
    return dec


def encoder_from_reference(module, device=None):
    """Drop-in encoder sharing the reference encoder's trunk (``module.resnet``) and copying its tail parameters."""
    from .encoders import EncoderCNN, EncoderCNNAtt
    if hasattr(module, "adaptive_pool"):
        size = module.adaptive_pool.output_size
        enc = EncoderCNNAtt(size[0] if isinstance(size, (tuple, list)) else size, backbone=module.resnet)
    else:
        enc = EncoderCNN(module.linear.out_features, backbone=module.resnet, in_features=module.linear.in_features)
        enc.linear.load_state_dict(module.linear.state_dict())
        enc.bn.load_state_dict(module.bn.state_dict())
        enc.bn.momentum = module.bn.momentum
    enc.train(module.training)
    return enc.to(device) if device is not None else enc


def optimizer_from_reference(opt, decoder, extra_params=()):
    """``FusedClampAdam`` carrying the moments / step counts / lr of a torch.optim.Adam over
    ``list(decoder.parameters()) + extra_params`` (the reference's parameter order, train_multitask.py:163-167)."""
    sd = opt.state_dict() if hasattr(opt, "state_dict") else opt
    g = sd["param_groups"][0]
    new = FusedClampAdam(decoder, lr=g["lr"], betas=tuple(g.get("betas", (0.9, 0.999))), eps=g.get("eps", 1e-8),
                         extra_params=extra_params)
    if len(g["params"]) != len(new.param_groups[0]["params"]):
        # e.g. exporter.py builds `optimizer` over decoder + encoder.adaptive_pool (no parameters): decoder part only
        n = len(list(decoder.parameters()))
        sd = {"state": {k: v for k, v in sd["state"].items() if int(k) < n},
              "param_groups": [dict(g, params=list(range(n)))]}
        new = FusedClampAdam(decoder, lr=g["lr"], betas=tuple(g.get("betas", (0.9, 0.999))), eps=g.get("eps", 1e-8))
    new.load_state_dict(sd)
    return new


def convert_checkpoint(ckpt, device=None):
    """Reference checkpoint (path or the un-pickled dict) -> the same dict with drop-in objects.  Un-pickling a
    whole-module checkpoint needs the reference's ``model`` / ``model_att`` module importable, exactly as in the
    reference itself (torch.load resolves the classes by name)."""
    if isinstance(ckpt, (str, os.PathLike)):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    out = dict(ckpt)
    dec = ckpt["decoder"]
    out["decoder"] = decoder_from_state_dict(dec, device=device) if isinstance(dec, dict) else decoder_from_reference(dec, device)
    extra = ()
    if ckpt.get("encoder") is not None and not isinstance(ckpt["encoder"], dict):
        out["encoder"] = encoder_from_reference(ckpt["encoder"], device)
        if hasattr(out["encoder"], "linear"):
            extra = list(out["encoder"].linear.parameters()) + list(out["encoder"].bn.parameters())
    for key, ex in (("optimizer", extra), ("lang_optimizer", ())):
        if ckpt.get(key) is not None:
            out[key] = optimizer_from_reference(ckpt[key], out["decoder"], ex)
    return out


def export_state(decoder, optimizer=None, lang_optimizer=None, encoder=None):
    """Plain tensors the reference classes load: ``decoder.load_state_dict(out['decoder'])``,
    ``torch.optim.Adam(...).load_state_dict(out['optimizer'])`` (and ``decoder-N.ckpt`` = ``out['decoder']``)."""
    out = {"decoder": {k: v.detach().cpu().clone() for k, v in decoder.state_dict().items()}}
    if encoder is not None:
        out["encoder"] = {k: v.detach().cpu().clone() for k, v in encoder.state_dict().items()}
    for key, opt in (("optimizer", optimizer), ("lang_optimizer", lang_optimizer)):
        if opt is not None:
            sd = opt.state_dict()
            sd["state"] = {k: {n: (t.detach().cpu() if torch.is_tensor(t) else t) for n, t in st.items()}
                           for k, st in sd["state"].items()}
            sd["param_groups"][0].pop("grad_clip", None)
            out[key] = sd
    return out


def save_checkpoint(folder, data_name, mode, epoch, epochs_since_improvement, encoder, decoder, optimizer,
                    lang_optimizer, bleu4, is_best):
    """Same arguments, dict layout and file names as stylenet/utils.py:63-90 (whole objects are pickled; the drop-in
    modules leave their per-process CUDA caches behind, see _DecoderBase.__getstate__)."""
    state = {"epoch": epoch, "epochs_since_improvement": epochs_since_improvement, "bleu-4": bleu4,
             "encoder": encoder, "decoder": decoder, "optimizer": optimizer, "lang_optimizer": lang_optimizer}
    filename = folder + "/" + mode + "_checkpoint_" + data_name + ".pth.tar"
    torch.save(state, filename)
    if is_best:
        torch.save(state, folder + "/" + mode + "_BEST_checkpoint_" + data_name + ".pth.tar")
    return filename


def load_model(checkpoint_path, device=None):
    """(encoder, decoder) in eval mode from a checkpoint written by the reference OR by ``save_checkpoint`` above
    (app/backend/model.py:490-499)."""
    ck = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    from .decoders import _DecoderBase
    if not isinstance(ck["decoder"], _DecoderBase):
        ck = convert_checkpoint(ck)
    enc, dec = ck.get("encoder"), ck["decoder"]
    if device is not None:
        dec = dec.to(device)
        enc = enc.to(device) if enc is not None else None
    dec.eval()
    if enc is not None:
        enc.eval()
    return enc, dec
