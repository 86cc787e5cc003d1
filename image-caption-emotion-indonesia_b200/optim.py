"""Fused clamp + Adam (K7): utils.clip_gradient (stylenet/utils.py:51-60) followed by
torch.optim.Adam.step (stylenet/train_multitask.py:388-389) as ONE kernel over the flat arena ranges that
received a gradient this step.  Several optimizer objects may share the same parameters, each with its own
moments and step counts (the reference keeps ``optimizer`` and ``lang_optimizer`` over the same decoder,
train_multitask.py:163-167); parameters whose grad is None are skipped exactly like torch.optim.Adam."""
import math

import torch

from . import ops


class FusedClampAdam:
    def __init__(self, decoder, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, grad_clip=0.5, extra_params=()):
        self.decoder = decoder
        self.lr, self.betas, self.eps, self.grad_clip = float(lr), betas, float(eps), float(grad_clip)
        self.param_groups = [{"lr": self.lr, "params": list(decoder.parameters()) + list(extra_params)}]
        self._arena_version = -1
        self.m = self.v = None
        self.steps = {}
        self.extra = list(extra_params)       # e.g. encoder.linear / encoder.bn parameters
        self.extra_state = {}

    def _state(self):
        a = self.decoder.arena()
        if self._arena_version != a.version:
            self.m = torch.zeros_like(a.flat)
            self.v = torch.zeros_like(a.flat)
            self.steps = {n: 0 for n in a.named}
            self._arena_version = a.version
        return a

    def zero_grad(self):
        for p in self.param_groups[0]["params"]:
            p.grad = None

    def _coeffs(self, step):
        b1, b2 = self.betas
        lr = self.param_groups[0]["lr"]
        return lr / (1.0 - b1 ** step), math.sqrt(1.0 - b2 ** step)

    @torch.no_grad()
    def step(self):
        a = self._state()
        items, foreign = a.grad_ranges()
        b1, b2 = self.betas
        ranges, ss, bc = [], [], []
        for off, n, name in items:
            self.steps[name] += 1
            s, c = self._coeffs(self.steps[name])
            if ranges and ranges[-1][0] + ranges[-1][1] == off and ss[-1] == s and bc[-1] == c:
                ranges[-1] = (ranges[-1][0], ranges[-1][1] + n)
            else:
                # arena groups are padded to 64 elements: merge across the (zero-gradient) padding too
                if ranges and off - (ranges[-1][0] + ranges[-1][1]) < 64 and ss[-1] == s and bc[-1] == c \
                        and self._gap_is_padding(a, ranges[-1][0] + ranges[-1][1], off):
                    ranges[-1] = (ranges[-1][0], off + n - ranges[-1][0])
                else:
                    ranges.append((off, n)); ss.append(s); bc.append(c)
        ops.adam_clamp(a.flat, a.gflat, self.m, self.v, ranges, ss, bc, b1, b2, self.eps, self.grad_clip)
        # parameters outside the arena (foreign grads or extra params): same kernel, one range each
        for name in foreign:
            self._step_tensor(a.named[name], ("arena", name))
        for i, p in enumerate(self.extra):
            if p.grad is not None:
                self._step_tensor(p, ("extra", i))

    @staticmethod
    def _gap_is_padding(a, lo, hi):
        # a gap is padding iff no parameter starts inside it
        return not any(lo <= o < hi for o in a.offset.values())

    def _step_tensor(self, p, key):
        st = self.extra_state.get(key)
        if st is None:
            st = {"m": torch.zeros_like(p.data), "v": torch.zeros_like(p.data), "step": 0}
            self.extra_state[key] = st
        st["step"] += 1
        s, c = self._coeffs(st["step"])
        g = p.grad.contiguous()
        pd = p.data if p.data.is_contiguous() else None
        if pd is None:
            raise RuntimeError("non-contiguous parameter")
        ops.adam_clamp(pd.view(-1), g.view(-1), st["m"].view(-1), st["v"].view(-1), [(0, p.numel())], [s], [c],
                       self.betas[0], self.betas[1], self.eps, self.grad_clip)
        if g.data_ptr() != p.grad.data_ptr():
            p.grad.copy_(g.view_as(p.grad))
