"""Host-side packing plan: what ``pack_padded_sequence(...).batch_sizes`` gives the reference
(stylenet/model.py:173-174) plus the row maps the kernels index with.  Pure host logic (numpy);
device copies are cached per (lengths, device)."""
from collections import OrderedDict

import numpy as np
import torch


def batch_sizes_from_lengths(lengths):
    """b_t = #{b: L_b > t}; ``lengths`` sorted descending, all > 0 (pack_padded_sequence contract)."""
    lengths = [int(l) for l in lengths]
    if len(lengths) == 0:
        raise ValueError("empty batch")
    if any(lengths[i] < lengths[i + 1] for i in range(len(lengths) - 1)):
        raise RuntimeError("`lengths` array must be sorted in decreasing order")
    if lengths[-1] <= 0:
        raise RuntimeError("Length of all samples has to be greater than 0")
    L = np.asarray(lengths)
    return [int((L > t).sum()) for t in range(lengths[0])]


class PackPlan:
    """Packed layout of one batch: row(b, t) = off[t] + b for b < bs[t]."""

    def __init__(self, lengths):
        self.lengths = [int(l) for l in lengths]
        self.B = len(self.lengths)
        self.bs = batch_sizes_from_lengths(self.lengths)
        self.T = len(self.bs)
        self.off = [0] * self.T
        for t in range(1, self.T):
            self.off[t] = self.off[t - 1] + self.bs[t - 1]
        self.N = self.off[-1] + self.bs[-1]
        row_b = np.concatenate([np.arange(b, dtype=np.int32) for b in self.bs])
        row_t = np.concatenate([np.full(b, t, dtype=np.int32) for t, b in enumerate(self.bs)])
        self.row_b_np, self.row_t_np = row_b, row_t
        self._dev = {}

    def dev(self, device):
        """Device copies: bs, off, row_b, row_t (int32) and flat (b*Tmax+t) index (int64)."""
        key = str(device)
        d = self._dev.get(key)
        if d is None:
            host = np.concatenate([np.asarray(self.bs, np.int32), np.asarray(self.off, np.int32),
                                   self.row_b_np, self.row_t_np])
            buf = torch.from_numpy(host).to(device)
            T, N = self.T, self.N
            d = {"bs": buf[:T], "off": buf[T:2 * T], "row_b": buf[2 * T:2 * T + N],
                 "row_t": buf[2 * T + N:2 * T + 2 * N]}
            d["flat_bt"] = d["row_b"].long() * T + d["row_t"].long()
            self._dev[key] = d
        return d


_plans = OrderedDict()


def get_plan(lengths):
    key = tuple(int(l) for l in lengths)
    p = _plans.get(key)
    if p is None:
        p = PackPlan(key)
        _plans[key] = p
        if len(_plans) > 256:
            _plans.popitem(last=False)
    else:
        _plans.move_to_end(key)
    return p


def shard_lengths(lengths, world_size, rank):
    """Data-parallel split AFTER the length sort: rank r takes samples r, r+W, r+2W, ... so every shard
    is itself sorted descending and shards are balanced in tokens."""
    idx = list(range(rank, len(lengths), world_size))
    return idx, [lengths[i] for i in idx]
