"""Oracle of the multi-layer FactoredLSTM stack (BASELINE.json configs[3]) -- TEST INFRASTRUCTURE ONLY.

The reference ignores ``num_layers`` (stylenet/model.py:37), so the stack is an EXTENSION whose definition is this
file: ``layers`` is a list of decoder objects exposing the reference's
``forward_step(embedded, states, mode)`` (stylenet/model.py:115-155) -- oracle-port decoders, or the UNMODIFIED
reference ``DecoderFactoredLSTM`` objects when /root/reference is present -- composed the way
``nn.LSTM(num_layers)`` composes cells (seq2seq/model.py:46-49):

    layer 0:  forward_step(x_t,            (h0, c0), mode)      embedding B / dropout / feature row of layers[0]
    layer l:  forward_step(h_t^{l-1},      (hl, cl), mode)      layers[l] built with embed_size = hidden_size
    logits:   layers[0].C(h_t^{top})

with the unroll, packing, teacher-forcing coin and arg-max feedback of stylenet/model.py:157-196 around it.
"""
import random

import torch

from oracle.port import batch_sizes_of


def stack_forward(layers, captions, lengths, features=None, teacher_forcing_ratio=0.8, mode="factual",
                  return_hidden=False):
    l0 = layers[0]
    emb = l0.dropout(l0.B(captions))                                     # model.py:166-167
    if features is not None:
        emb = torch.cat([features.unsqueeze(1), emb], 1)                 # model.py:170-171
    bs = batch_sizes_of(list(lengths))                                   # model.py:173-174
    Bn = captions.size(0)
    hs = [emb.new_zeros(Bn, l0.hidden_size) for _ in layers]             # model.py:176-177, per layer
    cs = [emb.new_zeros(Bn, l0.hidden_size) for _ in layers]
    tops = []
    predicted = captions[:, 0]                                           # model.py:179
    for t, b in enumerate(bs):
        if random.random() < teacher_forcing_ratio:                      # model.py:181
            x = emb[:b, t]
        else:
            x = l0.B(predicted)[:b]                                      # model.py:184
        for l, layer in enumerate(layers):
            x, (hs[l], cs[l]) = layer.forward_step(x, (hs[l][:b], cs[l][:b]), mode)   # model.py:185-186
        tops.append(x)
        predicted = l0.C(x).max(1)[1]                                    # model.py:189-191
    tops = torch.cat(tops, 0)
    out = l0.C(tops)                                                     # model.py:193-194
    return (out, tops) if return_hidden else out


def stack_parameters(layers):
    """Parameters the stack owns: everything of layer 0, and V/S/U/W of the upper layers (their B and C are unused)."""
    ps = list(layers[0].parameters())
    for layer in layers[1:]:
        ps += [p for n, p in layer.named_parameters() if not (n.startswith("B.") or n.startswith("C."))]
    return ps
