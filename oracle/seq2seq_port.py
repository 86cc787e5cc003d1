"""CPU restatement of the reference's seq2seq module (seq2seq/model.py) -- TEST INFRASTRUCTURE ONLY.

SURVEY.md section 8 f4: the greedy samplers and the TRUE multi-layer ``nn.LSTM`` (``num_layers`` is honoured here,
unlike in the StyleNet / NIC decoders), which gives the 3-layer stack of configs[3] a reference-defined relative.
The arithmetic of the reference lives in PyTorch (``nn.LSTM``, ``nn.Linear``, ``nn.Embedding``): the port keeps those
modules -- same parameter names, so state_dicts interchange -- and restates the control flow around them.

  RnnLM           the body shared by EncoderRNN (seq2seq/model.py:30-122) and DecoderRNN (:125-217)
  EncoderRNN      forward(features, src_tokens, lengths, tf) -> (logits [sum L, V], (h, c) [layers, b_last, H])
  DecoderRNN      forward(states, dst_tokens, lengths, tf)   -> logits   (the passed states are IGNORED: :169-172)
  Seq2Seq         encoder + one decoder per emotion (:220-301)
"""
import random

import torch
import torch.nn as nn

from oracle.port import batch_sizes_of


class RnnLM(nn.Module):
    def __init__(self, embed_size, hidden_size, vocab_size, num_layers, dropout=0.22, max_seq_length=40):
        super().__init__()
        self.max_seq_length, self.num_layers, self.hidden_size = max_seq_length, num_layers, hidden_size
        self.dropout = nn.Dropout(dropout)
        self.embed = nn.Embedding(vocab_size, embed_size)
        self.lstm = nn.LSTM(embed_size, hidden_size, num_layers, batch_first=True)
        self.linear = nn.Linear(hidden_size, vocab_size)

    def forward_step(self, embedded, states):
        """One step through all layers (model.py:52-67 / :146-161): states [layers, b, H] or (None, None) = zeros."""
        h, c = states
        b = embedded.size(0)
        zeros = lambda: embedded.new_zeros(self.num_layers, b, self.hidden_size)
        h = zeros() if h is None else h
        c = zeros() if c is None else c
        x = embedded.unsqueeze(1) if embedded.dim() == 2 else embedded
        out, (h, c) = self.lstm(x, (h, c))
        return out.squeeze(1), (h, c)

    def _unroll(self, rows, tokens, lengths, tf_ratio):
        """The time loop of model.py:78-97 / :173-192 over the padded input rows [B, T, E]."""
        bs = batch_sizes_of([int(l) for l in lengths])
        B = tokens.size(0)
        h = rows.new_zeros(self.num_layers, B, self.hidden_size)
        c = rows.new_zeros(self.num_layers, B, self.hidden_size)
        predicted = tokens[:, 0]
        hid = []
        for t, b in enumerate(bs):
            x = rows[:b, t] if random.random() < tf_ratio else self.embed(predicted)[:b]
            out, (h, c) = self.forward_step(x, (h[:, :b], c[:, :b]))
            hid.append(out)
            predicted = self.linear(out).max(1)[1]
        return self.linear(torch.cat(hid, 0)), (h, c)

    def _greedy(self, x, states):
        """model.py:99-122 / :194-217: max_seq_length arg-max steps, feeding the prediction back."""
        ids = []
        for _ in range(self.max_seq_length):
            out, states = self.forward_step(x, states)
            pred = self.linear(out).max(1)[1]
            ids.append(pred)
            x = self.embed(pred)
        return torch.stack(ids, 1), states


class EncoderRNN(RnnLM):
    def forward(self, features, src_tokens, lengths, teacher_forcing_ratio=0.5):
        rows = torch.cat([features.unsqueeze(1), self.dropout(self.embed(src_tokens))], 1)       # model.py:71-73
        return self._unroll(rows, src_tokens, lengths, teacher_forcing_ratio)

    def sample(self, features, states=(None, None)):
        return self._greedy(features, states)


class DecoderRNN(RnnLM):
    def forward(self, states, dst_tokens, lengths, teacher_forcing_ratio=0.5):
        rows = self.dropout(self.embed(dst_tokens))                                                # model.py:165-166
        return self._unroll(rows, dst_tokens, lengths, teacher_forcing_ratio)[0]                   # zero initial state

    def sample(self, start_token, states):
        x = self.embed(torch.tensor([start_token], dtype=torch.long))                              # batch of one
        return self._greedy(x, states)[0]


class Seq2Seq(nn.Module):
    def __init__(self, embed_size, hidden_size, vocab_size, num_layers, dropout=0.22, max_seq_length=40):
        super().__init__()
        self.hidden_size, self.max_seq_length = hidden_size, max_seq_length
        self.encoder = EncoderRNN(embed_size, hidden_size, vocab_size, num_layers, dropout=dropout)
        for s in ("happy", "sad", "angry"):
            setattr(self, "decoder_" + s, DecoderRNN(embed_size, hidden_size, vocab_size, num_layers, dropout=dropout))

    def forward(self, features, src, dst=(None, None), teacher_forcing_ratio=0.8, mode="factual"):
        outputs, states = self.encoder(features, src[0], src[1], teacher_forcing_ratio)
        if mode == "factual":
            return outputs
        return getattr(self, "decoder_" + mode)(states, dst[0], dst[1], teacher_forcing_ratio)

    def sample(self, features, start_token, states=(None, None), mode="factual"):
        ids, states = self.encoder.sample(features, states)
        if mode == "factual":
            return ids
        return getattr(self, "decoder_" + mode).sample(start_token, states)
