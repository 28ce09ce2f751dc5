"""Fused optimizer step: ``torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)`` +
``torch.optim.AdamW(param_dicts, lr, weight_decay).step()`` as the reference runs them every iteration (engine_vg.py:62-67,
main_vg.py:234-268), on device-flat buffers: all parameters are re-pointed into ONE flat fp32 buffer (state_dict, DDP and the
engine keep working on the same ``nn.Parameter`` objects), both moments are flat, and the gradients arrive flat from the engine
(``HotPathFunction`` hands autograd views of one buffer, which AccumulateGrad keeps without copying).  A step is then one
reduction (global norm) and one pass (clip + decoupled weight decay + Adam), with no host synchronisation.

Drop-in use (no reference file edited):
    optimizer = reftr_b200.optim.FusedAdamW(param_dicts, lr=args.lr, weight_decay=args.weight_decay)
    grad_total_norm = reftr_b200.optim.clip_grad_norm_(model.parameters(), max_norm)   # deferred into optimizer.step()
    optimizer.step()
LR schedulers (StepLR / LambdaLR, main_vg.py:269-287) work unchanged: ``param_groups[i]["lr"]`` is read on every step.

Semantics that differ from the torch pair, on purpose: (1) ``clip_grad_norm_`` is DEFERRED -- it returns the total norm, but ``p.grad``
itself stays unclipped; the coefficient is applied inside ``step()`` (read ``p.grad`` after clipping only through the optimizer).
(2) every optimized parameter must have a gradient at ``step()``: torch skips gradient-less parameters, a flat pass cannot, so that
case raises instead of silently applying weight decay to them.
"""
import weakref

import torch

from . import ops

_ACTIVE = weakref.WeakSet()
ALIGN = 64  # elements; the engine's flat gradient buffer uses the same slot alignment


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, layout=None):
        """``layout``: optional {parameter: element offset} (e.g. the engine's gradient slots, so that its flat gradient buffer is
        consumed without a copy); default = parameters in group order, ALIGN-element slots."""
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) > 8:
            raise ValueError("FusedAdamW supports up to 8 parameter groups (the reference uses 4, main_vg.py:234-262)")
        betas_set = {tuple(g["betas"]) for g in self.param_groups} | {g["eps"] for g in self.param_groups}
        if len(betas_set) != 2:
            raise ValueError("FusedAdamW: betas and eps must be the same for every group")
        self._plist = [(gi, p) for gi, g in enumerate(self.param_groups) for p in g["params"]]
        dev = {p.device for _, p in self._plist}
        if len(dev) != 1 or any(p.dtype != torch.float32 for _, p in self._plist):
            raise ValueError("FusedAdamW: all parameters must be fp32 on one device")
        self.device = dev.pop()
        if layout is None:
            layout, off = {}, 0
            for _, p in self._plist:
                layout[p] = off
                off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self._off = {id(p): int(layout[p]) for _, p in self._plist}
        order = sorted(self._plist, key=lambda gp: self._off[id(gp[1])])
        self.n = max(self._off[id(p)] + (p.numel() + ALIGN - 1) // ALIGN * ALIGN for _, p in order)
        # segments: maximal runs of consecutive slots that belong to the same group
        self.seg_end, self.seg_group = [], []
        for k, (gi, p) in enumerate(order):
            end = self._off[id(order[k + 1][1])] if k + 1 < len(order) else self.n
            if self.seg_group and self.seg_group[-1] == gi:
                self.seg_end[-1] = end
            else:
                self.seg_end.append(end)
                self.seg_group.append(gi)
        if len(self.seg_end) > 32:
            raise ValueError("FusedAdamW: parameter groups interleave in more than 32 runs; pass parameters in layout order")
        self.flat_p = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p)
        self.flat_g = None
        self._gviews = None
        with torch.no_grad():
            for _, p in self._plist:
                o = self._off[id(p)]
                view = self.flat_p[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view  # same nn.Parameter object, storage now inside the flat buffer
                self.state[p] = {"step": torch.zeros((), dtype=torch.float32),
                                 "exp_avg": self.flat_m[o:o + p.numel()].view(p.shape),
                                 "exp_avg_sq": self.flat_v[o:o + p.numel()].view(p.shape)}
        self._step = 0
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._pending = None  # (flat gradient, max_norm) set by clip_grad_norm_
        _ACTIVE.add(self)

    @classmethod
    def for_model(cls, model, params, **kw):
        """Uses the gradient-slot layout of the model's engine (reftr_b200/engine.py) when every optimized parameter has a slot,
        so that the engine's flat gradient buffer is consumed without a copy."""
        eng = model.engine() if hasattr(model, "engine") else None
        layout = None
        if eng is not None:
            slot = {id(p): eng.slots[n][0] for n, p in eng.named}
            groups = params if isinstance(params, (list, tuple)) and params and isinstance(params[0], dict) else [{"params": list(params)}]
            plist = [p for g in groups for p in g["params"]]
            if all(id(p) in slot for p in plist):
                layout = {p: slot[id(p)] for p in plist}
        return cls(params, layout=layout, **kw)

    # ------------------------------------------------------------------------------------------------------------
    def owns(self, params):
        return {id(p) for p in params} <= set(self._off)

    def _flat_grads(self):
        """The gradients as ONE flat buffer at the optimizer's offsets: zero-copy when they already are views of one storage at
        those offsets (the engine's flat gradient, which AccumulateGrad keeps without copying), else gathered with one foreach copy."""
        ps = [p for _, p in self._plist]
        g0 = ps[0].grad
        if g0 is not None and g0.dtype == torch.float32 and g0.is_contiguous():
            st = g0.untyped_storage()
            first = g0.storage_offset() - self._off[id(ps[0])]  # element offset of slot 0 inside that storage
            base = st.data_ptr()
            # every gradient must sit at its slot of that one storage (checked in full: ~0.3 ms of host time for 470 tensors)
            if first >= 0 and st.nbytes() >= 4 * (first + self.n) and (base + 4 * first) % 16 == 0 and all(
                    p.grad is not None and p.grad.data_ptr() == base + 4 * (first + self._off[id(p)]) and p.grad.is_contiguous()
                    and p.grad.dtype == torch.float32 for p in ps):
                return torch.empty(0, dtype=torch.float32, device=self.device).set_(st, first, (self.n,), (1,))
        return self._gather(ps)

    def _gather(self, ps):
        if self.flat_g is None:
            self.flat_g = torch.zeros(self.n, dtype=torch.float32, device=self.device)
            self._gviews = [self.flat_g[self._off[id(p)]:self._off[id(p)] + p.numel()].view(p.shape) for p in ps]
        missing = [p for p in ps if p.grad is None]
        if missing:
            # torch.optim.AdamW SKIPS a parameter without a gradient (no weight decay, no moment decay, its step counter stands
            # still); one flat pass cannot, and treating "no gradient" as a zero gradient would silently decay such weights.
            raise RuntimeError(f"FusedAdamW: {len(missing)} of {len(ps)} optimized parameters have no gradient (first shape "
                               f"{tuple(missing[0].shape)}); optimize only parameters that receive one every step (the engine hands a "
                               "gradient to every requires_grad parameter), or use torch.optim.AdamW")
        torch._foreach_copy_(self._gviews, [p.grad for p in ps])
        return self.flat_g

    def defer_clip(self, max_norm):
        """clip_grad_norm_ for this optimizer's parameters: computes the global norm now (one reduction, returned as a device
        scalar like torch's) and applies the clip coefficient inside the next ``step()``."""
        g = self._flat_grads()
        self._sumsq.zero_()
        ops.sumsq(g, self._sumsq)
        self._pending = (g, float(max_norm))
        return self._sumsq.sqrt().squeeze(0)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if self._pending is not None:
            g, max_norm = self._pending
            self._pending = None
            sumsq = self._sumsq
        else:
            g, max_norm, sumsq = self._flat_grads(), 0.0, None
        self._step += 1
        g0 = self.param_groups[0]
        ops.adamw_flat(self.flat_p, g, self.flat_m, self.flat_v, self.seg_end, self.seg_group,
                       [grp["lr"] for grp in self.param_groups], [grp["weight_decay"] for grp in self.param_groups],
                       g0["betas"][0], g0["betas"][1], g0["eps"], self._step, sumsq, max_norm)
        # the kernel wrote the parameters behind autograd's back: bump their version counters so that version-tracking consumers
        # (the engine's packed 16-bit weight copies, reftr_b200/pack.py) see the change, exactly as after torch's in-place update
        torch.autograd.graph.increment_version([p for _, p in self._plist])
        return loss

    def state_dict(self):
        for st in self.state.values():  # torch's per-parameter step counters are materialised only when somebody looks
            st["step"].fill_(self._step)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        with torch.no_grad():
            for _, p in self._plist:
                st, o = self.state[p], self._off[id(p)]
                for key, flat in (("exp_avg", self.flat_m), ("exp_avg_sq", self.flat_v)):
                    view = flat[o:o + p.numel()].view(p.shape)
                    view.copy_(st[key])
                    st[key] = view
        steps = [int(st["step"]) for st in self.state.values()]
        self._step = max(steps) if steps else 0
        for st in self.state.values():  # older checkpoints store the step as a python int
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32)


def clip_grad_norm_(parameters, max_norm, norm_type=2.0):
    """Drop-in for torch.nn.utils.clip_grad_norm_ (engine_vg.py:63).  When the parameters belong to a FusedAdamW the norm is one
    fused reduction and the scaling is deferred into its step(); otherwise this is torch's function."""
    params = [p for p in (parameters if not isinstance(parameters, torch.Tensor) else [parameters]) if p.grad is not None]
    if float(norm_type) == 2.0 and params:
        for opt in list(_ACTIVE):
            if opt.owns(params) and len(params) == len(opt._plist):
                return opt.defer_clip(max_norm)
    return torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type)
