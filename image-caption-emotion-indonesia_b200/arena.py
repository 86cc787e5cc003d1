"""Flat parameter / gradient arenas.

All decoder parameters live in ONE contiguous fp32 buffer in HBM (and their gradients in a second one),
laid out so that the gate-stacked matrices the kernels consume ([V_i;V_f;V_o;V_c], [W_i;..;W_c], the four
S of a style, ...) are contiguous sub-blocks.  ``nn.Parameter``s keep the reference names and shapes
(state_dicts interchange) but their storage is a view into the arena, so the fused clamp+Adam kernel and
the NCCL all-reduce each run over a handful of flat ranges instead of 59-89 tensors."""
import weakref

import torch

ALIGN = 64  # elements (256 B)

# Bumped whenever ANY module registers a parameter (``decoder.B.weight = nn.Parameter(pretrained)`` goes through
# Module.register_parameter): arenas compare it with the epoch they were bound at and re-validate every parameter
# OBJECT, not just the storage pointers of two sentinels, when it moved.
PARAM_EPOCH = [0]


def _on_register_parameter(module, name, param):
    PARAM_EPOCH[0] += 1
    return None


torch.nn.modules.module.register_module_parameter_registration_hook(_on_register_parameter)


class ParamArena:
    def __init__(self, module, groups):
        """``groups``: list of lists of parameter names; names inside a group are laid out back to back
        (no padding) so the group can be viewed as one stacked matrix."""
        self._module = weakref.ref(module)      # no decoder <-> arena reference cycle
        self.groups = groups
        self.named = dict(module.named_parameters())
        listed = [n for g in groups for n in g]
        missing = set(self.named) - set(listed)
        if missing or len(listed) != len(set(listed)):
            raise RuntimeError("arena layout does not cover parameters exactly once: %s" % sorted(missing))
        self.offset, self.numel = {}, {}
        self.group_span = []
        off = 0
        for g in groups:
            off = (off + ALIGN - 1) // ALIGN * ALIGN
            start = off
            for n in g:
                self.offset[n] = off
                self.numel[n] = self.named[n].numel()
                off += self.numel[n]
            self.group_span.append((start, off - start))
        self.total = (off + ALIGN - 1) // ALIGN * ALIGN
        self.flat = None
        self.gflat = None
        self.version = 0
        self.kernel_epoch = 0       # bumped whenever OUR kernels rewrite parameters (torch's ._version does not see them)
        self._sentinels = (listed[0], listed[-1])
        self._epoch = PARAM_EPOCH[0]

    # -- binding -------------------------------------------------------------------------------
    def bound(self):
        """Are the parameters still views into the arena?  ``.to()`` / ``.cuda()`` re-home every parameter at once, so
        the hot path only checks the first and the last one; every 256th call (and the first) checks all of them.
        (This runs several times per decode step: a full walk over 59-89 parameters each time was measurable.)"""
        if self.flat is None:
            return False
        if self._epoch != PARAM_EPOCH[0]:
            # some module registered a parameter since the last check: a Parameter object of this decoder may have
            # been replaced -> compare object identities, adopt the new objects and rebuild if anything changed
            self._epoch = PARAM_EPOCH[0]
            live = dict(self._module().named_parameters())
            if set(live) != set(self.named) or any(live[n] is not self.named[n] for n in live):
                if set(live) != set(self.named):
                    raise RuntimeError("the decoder's parameter set changed after the arena was laid out")
                self.named = live
                return False
        base = self.flat.data_ptr()
        self._checks = getattr(self, "_checks", 0) + 1
        if self._checks & 255 != 1:
            for n in self._sentinels:
                p = self.named[n]
                if p.data_ptr() != base + 4 * self.offset[n] or p.device != self.flat.device:
                    return False
            return True
        for n, p in self.named.items():
            if p.data_ptr() != base + 4 * self.offset[n] or p.device != self.flat.device:
                return False
        return True

    def ensure(self):
        """(Re)bind parameters into the arena (first use, or after .to()/.cuda() replaced storages)."""
        if self.bound():
            return self
        ref = next(iter(self.named.values()))
        for p in self.named.values():
            if p.dtype != torch.float32:
                raise RuntimeError("the decoder keeps fp32 master parameters; got %s" % p.dtype)
        flat = torch.zeros(self.total, dtype=torch.float32, device=ref.device)
        for n, p in self.named.items():
            view = flat[self.offset[n]:self.offset[n] + self.numel[n]].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self.flat = flat
        self.gflat = torch.zeros(self.total, dtype=torch.float32, device=ref.device)
        self.version += 1
        return self

    # -- views ---------------------------------------------------------------------------------
    def _span(self, names):
        o = self.offset[names[0]]
        n = 0
        for nm in names:
            if self.offset[nm] != o + n:
                raise RuntimeError("parameters %s are not contiguous in the arena" % (names,))
            n += self.numel[nm]
        return o, n

    def block(self, names, shape, grad=False):
        o, n = self._span(names)
        return (self.gflat if grad else self.flat)[o:o + n].view(shape)

    def grad_view(self, name):
        p = self.named[name]
        return self.gflat[self.offset[name]:self.offset[name] + self.numel[name]].view(p.shape)

    def is_arena_grad(self, name):
        p = self.named[name]
        return p.grad is not None and p.grad.data_ptr() == self.gflat.data_ptr() + 4 * self.offset[name]

    def publish_grads(self, names, fresh):
        """Expose freshly computed gradients of ``names`` as ``p.grad`` (arena views).  ``fresh`` is the
        flat buffer the kernels wrote (self.gflat, or a temporary when accumulating)."""
        for n in names:
            p = self.named[n]
            if not p.requires_grad:
                continue
            o, k = self.offset[n], self.numel[n]
            if fresh is self.gflat:
                p.grad = self.gflat[o:o + k].view(p.shape)
            else:
                if p.grad is None:
                    gv = self.gflat[o:o + k].view(p.shape)
                    gv.copy_(fresh[o:o + k].view(p.shape))
                    p.grad = gv
                else:
                    p.grad.add_(fresh[o:o + k].view(p.shape))

    def content_key(self):
        """Changes whenever the parameter VALUES may have changed: in-place torch ops bump the version counter of the
        Parameter they touch (or of the flat buffer), the optimizer kernels / graph replays bump ``kernel_epoch``, a
        re-bind bumps ``version``."""
        return (self.version, self.kernel_epoch, self.flat._version, sum(p._version for p in self.named.values()))

    def grad_ranges(self):
        """Merged flat (offset, length) ranges of parameters whose .grad is an arena view, plus the list
        of parameters with foreign gradients."""
        items = []
        foreign = []
        for n, p in self.named.items():
            if p.grad is None:
                continue
            if self.is_arena_grad(n):
                items.append((self.offset[n], self.numel[n], n))
            else:
                foreign.append(n)
        items.sort()
        return items, foreign
