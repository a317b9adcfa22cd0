"""`FusedAdam`: torch.optim.Adam semantics (train_lat_celebhq_unet_cont2.py:477, lr 5e-5; train_poly_unet_cont.py:782,
lr 5e-4; betas (0.9, 0.999), eps 1e-8, no weight decay / amsgrad) as ONE kernel launch over all parameter tensors
(`sbm_adam_step`), instead of the reference's per-tensor loop over 286-354 tensors.

`GraphedTrainStep`: the reference's whole DSM training step (train_lat_celebhq_unet_cont2.py:95-100: loss_fn ->
zero_grad -> backward -> Adam.step) captured ONCE as a CUDA graph and replayed per batch.  A Poly-sized step is ~700
kernels of a few microseconds each, i.e. launch-bound from Python; everything that changes between steps (Philox draw
id, Adam step count) lives in device memory and is advanced by a 1-thread tick kernel inside the graph.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_CHUNK = 65536


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tables = {}
        self._stage = {}  # per group: pinned staging + device tables (allocated once, refreshed in place)
        self.capturable = capturable
        self.step_dev = None  # device counter of completed steps (capturable mode)

    def _table(self, gi, group):
        """Device-side descriptor + chunk tables; rebuilt whenever a grad / state pointer changes."""
        ps = [p for p in group["params"] if p.grad is not None]
        for p in ps:
            st = self.state[p]
            if "step" in st and not isinstance(st["step"], int):
                st["step"] = int(st["step"])  # a torch.optim.Adam checkpoint stores the step as a float tensor
            if not st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise L.SbmError("FusedAdam needs contiguous fp32 CUDA parameters")
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
        sig = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p in ps)
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
        dev = ps[0].device
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        cht = torch.tensor(chunks, dtype=torch.int32)
        stage = self._stage.get(gi)
        if stage is None or stage[0].numel() != raw.numel() or stage[1].shape != cht.shape:
            # pinned staging + device tables are allocated once (outside any CUDA-graph capture) and refreshed in place:
            # a captured copy node re-reads the pinned buffer at every replay
            stage = (raw.pin_memory(), cht.pin_memory(), torch.empty(raw.numel(), dtype=torch.uint8, device=dev),
                     torch.empty(cht.shape, dtype=torch.int32, device=dev))
            self._stage[gi] = stage
        else:
            stage[0].copy_(raw)
            stage[1].copy_(cht)
        stage[2].copy_(stage[0], non_blocking=True)
        stage[3].copy_(stage[1], non_blocking=True)
        self._tables[gi] = (sig, stage[2], stage[3], len(chunks), ps)
        return stage[2], stage[3], len(chunks), ps

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        stepped = False
        for gi, group in enumerate(self.param_groups):
            if not any(p.grad is not None for p in group["params"]):
                continue
            t_dev, c_dev, n_chunks, ps = self._table(gi, group)
            step = self.state[ps[0]]["step"] + 1
            for p in ps:
                self.state[p]["step"] = step
            if self.capturable and self.step_dev is None:
                self.step_dev = torch.full((1,), step - 1, dtype=torch.int32, device=ps[0].device)
            stepped = True
            b1, b2 = group["betas"]
            L.check(L.lib().sbm_adam_step(L.ptr(t_dev), L.ptr(c_dev), C.c_int32(n_chunks), C.c_int32(_CHUNK),
                                          C.c_float(group["lr"]), C.c_float(b1), C.c_float(b2), C.c_float(group["eps"]),
                                          C.c_int32(step), C.c_float(grad_scale),
                                          L.ptr(self.step_dev) if self.capturable else None, L.stream_ptr()),
                    "sbm_adam_step")
            # the kernel wrote the parameters through raw pointers: tell autograd (and the score net's packed-weight
            # cache, keyed by `_version`) that they changed
            torch.autograd.graph.increment_version(ps)
        if self.capturable and stepped:
            # the device counter is what the kernel's bias correction reads: advance it HERE (eager loops and captured
            # graphs alike), once per step whatever the number of parameter groups
            L.check(L.lib().sbm_train_tick(L.ptr(self.step_dev), None, C.c_uint64(0), L.stream_ptr()), "sbm_train_tick")
        return loss

    def _sync_host_steps(self):
        """Graph replays advance only the device counter: bring the host-side `state[p]['step']` up to date."""
        if self.capturable and self.step_dev is not None:
            n = int(self.step_dev.item())
            for st in self.state.values():
                if "step" in st:
                    st["step"] = n

    def state_dict(self):
        self._sync_host_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # the moment tensors were replaced: drop the raw-pointer tables that still point at the old ones, and restart
        # the device step counter from the loaded step
        self._tables.clear()
        step = 0
        for st in self.state.values():
            if "step" in st:
                st["step"] = int(st["step"])
                step = max(step, st["step"])
            for k in ("exp_avg", "exp_avg_sq"):
                if k in st:
                    st[k] = st[k].float().contiguous()
        if self.capturable and self.step_dev is not None:
            self.step_dev.fill_(step)


class GraphedTrainStep:
    """One DSM training step as a replayable CUDA graph.

        step = GraphedTrainStep(model, sde, example_batch, lr=5e-4)     # warm-up (3 eager steps) + capture
        for batch in loader: loss = step(batch)                          # copy-in + one graph launch; loss: 0-d tensor

    Semantics = the eager sequence `loss = loss_fn(batch, model, sde, ...); opt.zero_grad(); loss.backward();
    opt.step()` with in-kernel Philox (t, z) draws; verified step-for-step against it (tests/test_backward_gpu.py).
    `model` may be a `DataParallelScoreNet`: the bucketed NCCL all-reduces are captured with the step (verified on 2
    GPUs).  In that case call `close()` (or drop the object) BEFORE `dist.destroy_process_group()`: tearing the
    communicator down while a graph still holds its kernels dead-locks at exit."""

    def __init__(self, model, sde, example_batch, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, loss_kwargs=None, warmup=3,
                 optimizer=None):
        from . import sde_helper2 as sh
        self.model, self.sde = model, sde
        self.loss_kwargs = dict(reduce_mean=True, likelihood_weighting=False, eps=1e-5)
        self.loss_kwargs.update(loss_kwargs or {})
        self.opt = optimizer or FusedAdam(model.parameters(), lr=lr, betas=betas, eps=eps, capturable=True)
        if not self.opt.capturable:
            raise L.SbmError("GraphedTrainStep needs FusedAdam(capturable=True)")
        dev = example_batch.device
        self.batch = example_batch.detach().clone().float().contiguous()
        self.draw_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._sh = sh
        self._params = [p for p in model.parameters() if p.requires_grad]
        self._base_draw = sh._rng.draw  # host draw id baked into the graph; the device offset advances it
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            # >= 2 eager steps: the second one builds the device tables (multi-tensor re-pack, Adam) the capture re-uses
            for _ in range(max(warmup, 2)):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        n0 = L.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        # thread_local: other threads (e.g. the NCCL watchdog polling events) may touch CUDA while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss = self._body()
        self.launches_per_step = L.launch_count() - n0
        # capture does not execute: the counters still hold the post-warm-up state the first replay must start from

    def _body(self):
        sh = self._sh
        sh._rng.draw = self._base_draw  # same baked id every time; the device-side offset makes the draws differ
        loss = sh.loss_fn(self.batch, self.model, self.sde, rng="philox", draw_dev=self.draw_dev, **self.loss_kwargs)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        # (the Adam step count is advanced by FusedAdam.step itself)
        L.check(L.lib().sbm_train_tick(None, L.ptr(self.draw_dev), C.c_uint64(2), L.stream_ptr()), "sbm_train_tick")
        return loss

    def close(self):
        """Release the captured graph (required before destroying the process group when it holds NCCL kernels)."""
        self.graph = None
        self.loss = None
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def __call__(self, batch):
        self.batch.copy_(batch, non_blocking=True)
        self.graph.replay()
        # replays update the parameters behind autograd's back: invalidate version-keyed caches for later eager use
        torch.autograd.graph.increment_version(self._params)
        return self.loss


_ema_tables: dict = {}


@torch.no_grad()
def update_ema(ema_model, model, decay=0.999):
    """utils.py:79-90 (`ema_params[name].mul_(decay).add_(param, alpha=1-decay)` over every parameter;
    train_lat_celebhq_unet_cont2_cond.py:129, 672-674) as ONE multi-tensor kernel launch."""
    ema_params = dict(ema_model.named_parameters())
    pairs = [(ema_params[name.replace("module.", "")], p) for name, p in model.named_parameters()]
    sig = tuple((e.data_ptr(), p.data_ptr()) for e, p in pairs)
    key = (id(ema_model), id(model))
    hit = _ema_tables.get(key)
    if hit is None or hit[0] != sig:
        arr = (L.EmaTensor * len(pairs))()
        chunks = []
        for i, (e, p) in enumerate(pairs):
            if not (e.is_cuda and p.is_cuda and e.dtype == p.dtype == torch.float32 and e.is_contiguous()
                    and p.is_contiguous() and e.numel() == p.numel()):
                raise L.SbmError("update_ema needs matching contiguous fp32 CUDA parameters")
            arr[i] = L.EmaTensor(e.data_ptr(), p.data_ptr(), p.numel())
            chunks += [(i, k) for k in range((p.numel() + _CHUNK - 1) // _CHUNK)]
        dev = pairs[0][0].device
        hit = (sig, torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev),
               torch.tensor(chunks, dtype=torch.int32).to(dev), len(chunks))
        _ema_tables[key] = hit
    L.check(L.lib().sbm_ema_step(L.ptr(hit[1]), L.ptr(hit[2]), C.c_int32(hit[3]), C.c_int32(_CHUNK), C.c_float(decay),
                                 L.stream_ptr()), "sbm_ema_step")
    torch.autograd.graph.increment_version([e for e, _ in pairs])
