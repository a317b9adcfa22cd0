"""`FusedAdam`: torch.optim.Adam semantics (train_lat_celebhq_unet_cont2.py:477, lr 5e-5; train_poly_unet_cont.py:782,
lr 5e-4; betas (0.9, 0.999), eps 1e-8, no weight decay / amsgrad) as ONE kernel launch over all parameter tensors
(`sbm_adam_step`), instead of the reference's per-tensor loop over 286-354 tensors.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_CHUNK = 65536


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tables = {}

    def _table(self, gi, group):
        """Device-side descriptor + chunk tables; rebuilt whenever a grad / state pointer changes."""
        ps = [p for p in group["params"] if p.grad is not None]
        for p in ps:
            st = self.state[p]
            if not st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise L.SbmError("FusedAdam needs contiguous fp32 CUDA parameters")
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
        sig = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == sig:
            return hit[1:]
        arr = (L.AdamTensor * len(ps))()
        chunks = []
        for i, p in enumerate(ps):
            st = self.state[p]
            arr[i] = L.AdamTensor(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                  p.numel())
            chunks += [(i, k) for k in range((p.numel() + _CHUNK - 1) // _CHUNK)]
        raw = bytes(arr)
        dev = ps[0].device
        t_dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        c_dev = torch.tensor(chunks, dtype=torch.int32).to(dev)
        self._tables[gi] = (sig, t_dev, c_dev, len(chunks), ps)
        return t_dev, c_dev, len(chunks), ps

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if not any(p.grad is not None for p in group["params"]):
                continue
            t_dev, c_dev, n_chunks, ps = self._table(gi, group)
            step = self.state[ps[0]]["step"] + 1
            for p in ps:
                self.state[p]["step"] = step
            b1, b2 = group["betas"]
            L.check(L.lib().sbm_adam_step(L.ptr(t_dev), L.ptr(c_dev), C.c_int32(n_chunks), C.c_int32(_CHUNK),
                                          C.c_float(group["lr"]), C.c_float(b1), C.c_float(b2), C.c_float(group["eps"]),
                                          C.c_int32(step), C.c_float(grad_scale), L.stream_ptr()), "sbm_adam_step")
            # the kernel wrote the parameters through raw pointers: tell autograd (and the score net's packed-weight
            # cache, keyed by `_version`) that they changed
            torch.autograd.graph.increment_version(ps)
        return loss
