"""seq2seq modules of the reference (seq2seq/model.py) on the kernels -- SURVEY.md section 8 f4: the greedy samplers
and the TRUE multi-layer ``nn.LSTM`` (``num_layers`` is honoured here, unlike in the StyleNet / NIC decoders).

  EncoderRNN(embed_size, hidden_size, vocab_size, num_layers, dropout=0.22, max_seq_length=40)   seq2seq/model.py:30-122
      forward(features, src_tokens, lengths, teacher_forcing_ratio=0.5) -> (logits [sum L, V], (h, c) [layers, b_last, H])
      forward_step(embedded, (h, c)) ; sample(features, states=(None, None)) -> (ids [B, max_seq_length], states)
  DecoderRNN (seq2seq)                                                                            :125-217
      forward(states, dst_tokens, lengths, teacher_forcing_ratio=0.5) -> logits  (``states`` is ignored, as in :169-172)
      sample(start_token, states) -> ids [1, max_seq_length]
  Seq2Seq                                                                                         :220-301

Parameter names are ``nn.LSTM``'s (``lstm.weight_ih_l{k}`` ...), so state_dicts interchange with the reference.  The
layers reuse K2 (time-batched input projection of every layer), K3 (recurrence; layer l reads layer l-1's h rows, the
nn.LSTM stacking), K5 and the few-row decode kernels unchanged; the multi-layer plumbing mirrors stack.py."""
import torch
import torch.nn as nn

from . import ops
from .decoders import _Ctx, _DecoderBase, _LogitsFn


class _RnnLM(_DecoderBase):
    cell = ops.CELL_LSTM
    _layered = True          # decode keeps one (h, c) per layer

    def __init__(self, embed_size, hidden_size, vocab_size, num_layers, dropout=0.22, max_seq_length=40):
        super().__init__()
        self.max_seq_length, self.num_layers, self.hidden_size = max_seq_length, int(num_layers), hidden_size
        self.embed_size, self.vocab_size = embed_size, vocab_size
        self.dropout = nn.Dropout(dropout)
        self.embed = nn.Embedding(vocab_size, embed_size)
        self.lstm = nn.LSTM(embed_size, hidden_size, num_layers, batch_first=True)     # parameter container
        self.linear = nn.Linear(hidden_size, vocab_size)

    # -- layout ------------------------------------------------------------------------------------
    def _arena_groups(self):
        groups = [["embed.weight"]]
        for l in range(self.num_layers):
            groups += [["lstm.weight_ih_l%d" % l], ["lstm.weight_hh_l%d" % l], ["lstm.bias_ih_l%d" % l],
                       ["lstm.bias_hh_l%d" % l]]
        return groups + [["linear.weight"], ["linear.bias"]]

    def _emb(self):
        return self.embed

    def _emb_name(self):
        return "embed.weight"

    def _out(self):
        return self.linear

    def _out_names(self):
        return ("linear.weight", "linear.bias")

    def _p(self, kind, layer):
        self.arena()
        return getattr(self.lstm, "%s_l%d" % (kind, layer))

    def _recurrent_weights(self, layer=0):
        return self._p("weight_hh", layer), self._p("bias_hh", layer)

    def _recurrent_grads(self, gbuf, layer=0):
        H = self.hidden_size
        return (self._gview(gbuf, ["lstm.weight_hh_l%d" % layer], (4 * H, H)),
                self._gview(gbuf, ["lstm.bias_hh_l%d" % layer], (4 * H,)))

    def _seq_grad_names(self, mode):
        names = ["embed.weight"]
        for l in range(self.num_layers):
            names += ["lstm.%s_l%d" % (k, l) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        return names

    # -- K2: x W_ih^T + b_ih of one layer --------------------------------------------------------------
    def _input_projection(self, c, X, mode, r0, n):
        H = self.hidden_size
        L = getattr(c, "layer", 0)
        W, b = self._p("weight_ih", L), self._p("bias_ih", L)
        Ein = W.shape[1]
        if self.bf16:
            w16 = c.__dict__.setdefault("w16", {})
            if "Wih" not in w16:
                w16["Wih"] = self._shadow(W)
            Wb = w16["Wih"]
            Ep = Wb.stride(0)
            if X is not None:
                if c.__dict__.get("Xb") is None or (r0 == 0 and n == X.shape[0]):
                    c.Xb = torch.empty(X.shape[0], Ep, dtype=torch.bfloat16, device=X.device)
                ops.cast_bf16(X, n, Ein, Ein, c.Xb, Ep, Ep, src_off=r0 * Ein, dst_off=r0 * Ep)
            ops.gemm_bf16(ops.OP_NT, c.Xb, Wb, n, 4 * H, Ep, c.Xb.stride(0), Ep, C=c.XP, ldc=4 * H, bias=b,
                          a_off=r0 * c.Xb.stride(0), c_off=r0 * 4 * H)
            return
        ops.gemm(ops.OP_NT, X, W, c.XP, n, 4 * H, Ein, Ein, Ein, 4 * H, bias=b, a_off=r0 * Ein, c_off=r0 * 4 * H)

    def _input_projection_bwd(self, c, dZ, gbuf):
        H = self.hidden_size
        L = getattr(c, "layer", 0)
        W = self._p("weight_ih", L)
        Ein = W.shape[1]
        N = dZ.shape[0]
        gW = self._gview(gbuf, ["lstm.weight_ih_l%d" % L], (4 * H, Ein))
        gb = self._gview(gbuf, ["lstm.bias_ih_l%d" % L], (4 * H,))
        dX = torch.empty(N, Ein, dtype=torch.float32, device=dZ.device)
        if self.bf16:
            Wb = c.w16["Wih"]
            ops.gemm_bf16(ops.OP_TN, c.dZb, c.Xb, 4 * H, Ein, N, 4 * H, c.Xb.stride(0), C=gW, ldc=Ein)
            ops.colsum(dZ, N, 4 * H, 4 * H, gb)
            ops.gemm_bf16(ops.OP_NN, c.dZb, Wb, N, Ein, 4 * H, 4 * H, Wb.stride(0), C=dX, ldc=Ein)
            return dX
        ops.gemm(ops.OP_TN, dZ, c.X, gW, 4 * H, Ein, N, 4 * H, Ein, Ein)
        ops.colsum(dZ, N, 4 * H, 4 * H, gb)
        ops.gemm(ops.OP_NN, dZ, W, dX, N, Ein, 4 * H, 4 * H, Ein, Ein)
        return dX

    # -- layers above the first (the nn.LSTM stacking, seq2seq/model.py:46-49) -------------------------
    def _upper_layers_init(self, c, save):
        plan = c.plan
        dev = c.XP.device
        H, N, B = self.hidden_size, plan.N, plan.B
        f32 = dict(dtype=torch.float32, device=dev)
        b16 = dict(dtype=torch.bfloat16, device=dev)
        use_tc = self.bf16 and H % 32 == 0
        upper = []
        for l in range(1, self.num_layers):
            cl = _Ctx()
            cl.layer, cl.mode, cl.plan, cl.Ein = l, c.mode, plan, H
            cl.w16 = {}
            cl.X = cl.Xb = None
            cl.XP = torch.empty(N, 4 * H, **f32)
            cl.Hall = torch.empty(N, H, **f32)
            cl.Call = torch.empty(N, H, **f32) if save else None
            cl.Hprev = torch.empty(N, H, **f32) if (save and not use_tc) else None
            cl.gates = torch.empty(N, 4 * H, **f32) if save else None
            cl.c_state = torch.zeros(B, H, **f32)
            cl.Hb = cl.Hpb = None
            cl.Whh, cl.bhh = self._recurrent_weights(l)
            if use_tc:
                cl.w16["Whh"] = self._shadow(cl.Whh)
                cl.Hb = torch.empty(N, H, **b16)
                cl.Hpb = torch.empty(N, H, **b16) if save else None
            upper.append(cl)
        return upper

    def _upper_layers_fwd(self, c, t0, t1):
        plan = c.plan
        r0 = plan.off[t0]
        n = (plan.off[t1] if t1 < plan.T else plan.N) - r0
        below = c
        for cl in c.upper:
            if self.bf16 and below.Hb is not None:
                cl.Xb, X = below.Hb, None            # the recurrence below already wrote its h_t as the bf16 operand
            else:
                cl.X = X = below.Hall
            self._input_projection(cl, X, c.mode, r0, n)
            self._recur_fwd(c, cl, t0, t1)
            below = cl

    def _run_forward(self, plan, captions, features, coins, mode, save):
        c = super()._run_forward(plan, captions, features, coins, mode, save)
        self.__dict__["_last_ctx"] = c               # forward() reads the final states of every layer from it
        return c

    def _final_states(self):
        """(h, c) [layers, b_last, H] after the last step -- what the reference's loop leaves in h_t / c_t
        (seq2seq/model.py:88-97: the states are sliced to the batch size of every step, so only the samples alive at
        the last step remain)."""
        c = self.__dict__.pop("_last_ctx")
        plan = c.plan
        b, r0 = plan.bs[-1], plan.off[-1]
        ctxs = [c] + list(c.upper)
        h = torch.stack([x.Hall[r0:r0 + b] for x in ctxs], 0)
        cc = torch.stack([x.c_state[:b] for x in ctxs], 0)
        return h, cc

    # -- decode ----------------------------------------------------------------------------------------
    def _small_step(self, ctx, X, mode, R, h_prev, c_prev, src_row, h_out, c_out, layer=0, x_rows=None):
        H = self.hidden_size
        W = self._p("weight_ih", layer)
        if x_rows is not None:
            X = X.index_select(0, x_rows.long())
        ops.decode_cell(self.cell, H, R, W, W.shape[1], X, 0, self._p("bias_ih", layer), self._p("weight_hh", layer),
                        self._p("bias_hh", layer), h_prev, c_prev, src_row, h_out, c_out)

    def forward_step(self, embedded, states):
        """One step through all layers (seq2seq/model.py:52-67): states [layers, b, H] each, or (None, None) = zeros."""
        from .decode import single_step
        h, c = states
        x = embedded.reshape(embedded.shape[0], -1) if embedded.dim() == 3 else embedded
        b = x.shape[0]
        z = lambda: torch.zeros(self.num_layers, b, self.hidden_size, dtype=torch.float32, device=x.device)
        h = z() if h is None else h
        c = z() if c is None else c
        return single_step(self, x, (h, c), None)

    @torch.no_grad()
    def _greedy(self, x, states):
        """max_seq_length arg-max steps feeding the prediction back (seq2seq/model.py:99-122 / :194-217)."""
        ids = []
        V = self.vocab_size
        for _ in range(self.max_seq_length):
            out, states = self.forward_step(x, states)
            R = out.shape[0]
            logits = torch.empty(R, V, dtype=torch.float32, device=out.device)
            if R <= ops.SKINNY_MAX_ROWS:
                ops.skinny_linear(self.linear.weight, out, logits, R, bias=self.linear.bias)
            else:
                ops.gemm(ops.OP_NT, out, self.linear.weight, logits, R, V, self.hidden_size, self.hidden_size,
                         self.hidden_size, V, bias=self.linear.bias)
            am = torch.empty(R, dtype=torch.int64, device=out.device)
            ops.softmax_nll(logits, R, V, argmax=am)               # lowest index on ties, like torch.max
            ids.append(am)
            x = torch.empty(R, self.embed_size, dtype=torch.float32, device=out.device)
            ops.gather_pack_fwd(am.view(R, 1), self.embed.weight, None, False, _arange32(R, out.device),
                                _zeros32(R, out.device), None, R, x, 0.0, 0)
        return torch.stack(ids, 1), states


_small = {}


def _arange32(n, dev):
    k = ("a", n, str(dev))
    if k not in _small:
        _small[k] = torch.arange(n, dtype=torch.int32, device=dev)
    return _small[k]


def _zeros32(n, dev):
    k = ("z", n, str(dev))
    if k not in _small:
        _small[k] = torch.zeros(n, dtype=torch.int32, device=dev)
    return _small[k]


class EncoderRNN(_RnnLM):
    """seq2seq/model.py:30-122."""

    def forward(self, features, src_tokens, lengths, teacher_forcing_ratio=0.5):
        hall, _ = self._forward_hidden(src_tokens, lengths, features, teacher_forcing_ratio, None)
        states = self._final_states()
        return _LogitsFn.apply(hall, self.linear.weight, self), states

    def sample(self, features, states=(None, None)):
        self._check_inputs(features, features)
        return self._greedy(features.detach().float().contiguous(), states)


class DecoderRNN(_RnnLM):
    """seq2seq/model.py:125-217 (class name as in the reference; the NIC decoder of the same name lives in decoders.py)."""

    def forward(self, states, dst_tokens, lengths, teacher_forcing_ratio=0.5):
        hall, _ = self._forward_hidden(dst_tokens, lengths, None, teacher_forcing_ratio, None)   # zero initial state (:169-172)
        self.__dict__.pop("_last_ctx", None)
        return _LogitsFn.apply(hall, self.linear.weight, self)

    def sample(self, start_token, states):
        dev = self.embed.weight.device
        self._check_inputs(self.embed.weight, None)
        x = self.embed.weight[int(start_token)].detach().reshape(1, -1).contiguous()
        h, c = states
        return self._greedy(x, (h.to(dev).float().contiguous(), c.to(dev).float().contiguous()))[0]


class Seq2Seq(nn.Module):
    """seq2seq/model.py:220-301: one encoder (the factual captioner) and one decoder per emotion."""

    def __init__(self, embed_size, hidden_size, vocab_size, num_layers, dropout=0.22, max_seq_length=40):
        super().__init__()
        self.hidden_size, self.max_seq_length = hidden_size, max_seq_length
        self.encoder = EncoderRNN(embed_size, hidden_size, vocab_size, num_layers, dropout=dropout)
        self.decoder_happy = DecoderRNN(embed_size, hidden_size, vocab_size, num_layers, dropout=dropout)
        self.decoder_sad = DecoderRNN(embed_size, hidden_size, vocab_size, num_layers, dropout=dropout)
        self.decoder_angry = DecoderRNN(embed_size, hidden_size, vocab_size, num_layers, dropout=dropout)

    def set_precision(self, precision):
        for m in (self.encoder, self.decoder_happy, self.decoder_sad, self.decoder_angry):
            m.set_precision(precision)
        return self

    def _decoder(self, mode):
        if mode not in ("happy", "sad", "angry"):
            raise ValueError("mode name wrong: %r" % (mode,))
        return getattr(self, "decoder_" + mode)

    def forward(self, features, src, dst=(None, None), teacher_forcing_ratio=0.8, mode="factual"):
        outputs, states = self.encoder(features, src[0], src[1], teacher_forcing_ratio)
        if mode == "factual":
            return outputs
        return self._decoder(mode)(states, dst[0], dst[1], teacher_forcing_ratio)

    def sample(self, features, start_token, states=(None, None), mode="factual"):
        ids, states = self.encoder.sample(features, states)
        if mode == "factual":
            return ids
        return self._decoder(mode).sample(start_token, states)
