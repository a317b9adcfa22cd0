"""Multi-GPU plumbing of the path (one process per GPU, `torch.distributed` over NCCL / NVLink).

The reference is single-GPU (`--cuda N`, train_lat_celebhq_unet_cont2.py:402); two things shard:

* **Sampling** shards over the batch: weights replicated, every rank owns `B/G` latents, no activation exchange.
  `shard_range` + `seed_shard` give each rank the slice of the global Philox stream its samples would have drawn
  in the unsharded batch, `corrector_allreduce` is the optional exact mode (the corrector's step size couples the
  batch through two means, sde_helper2.py:97-99: an all-reduce of two doubles between the norms and the update
  kernel), `gather_batch` is the final all-gather of the `[B/G, M, D, D]` shards.
* **DSM training** is data parallel: `DataParallelScoreNet` averages the gradients over the ranks with bucketed
  all-reduces that are issued from INSIDE the hand-written backward pass as soon as a bucket's last gradient has been
  produced (NCCL runs them on its own stream, so they overlap the remaining backward kernels); the identical
  `FusedAdam` step then runs on every rank.

Everything here is host logic over `torch.distributed`; it runs unchanged on the `gloo` backend with CPU tensors
(tests/test_distributed_cpu.py, world_size 2).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def shard_range(global_batch: int, rank: int | None = None, world: int | None = None, group=None):
    """[lo, hi) of the global batch owned by `rank`: contiguous, sizes differ by at most one, earlier ranks larger."""
    w, r = _world(group)
    world = w if world is None else world
    rank = r if rank is None else rank
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def seed_shard(seed: int, global_batch: int, group=None):
    """Seed the in-kernel Philox stream so this rank draws what samples [lo, hi) of the unsharded batch would."""
    from . import sde_helper2 as sh
    lo, hi = shard_range(global_batch, group=group)
    sh.manual_seed(seed, sample_offset=lo)
    return lo, hi


def corrector_allreduce(group=None):
    """`reduce_fn` for `corrector` / `pc_sampler` / `cond_sampler`: sums the two batch-norm accumulators over the ranks
    (pass `global_batch=` too), so a sharded run takes exactly the step sizes of the unsharded batch."""
    def reduce_fn(acc2: torch.Tensor):
        if _world(group)[0] > 1:
            dist.all_reduce(acc2, op=dist.ReduceOp.SUM, group=group)
    return reduce_fn


def gather_batch(x_local: torch.Tensor, global_batch: int | None = None, group=None) -> torch.Tensor:
    """Concatenate the ranks' batch shards (the one collective of sharded sampling).  Shards may differ by one sample
    (see `shard_range`); pass `global_batch` in that case."""
    world, _ = _world(group)
    if world == 1:
        return x_local
    if global_batch is None or global_batch % world == 0:
        out = torch.empty((x_local.shape[0] * world, *x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        dist.all_gather_into_tensor(out, x_local.contiguous(), group=group)
        return out
    # ragged shards: pad every shard to the largest, gather, trim
    sizes = [shard_range(global_batch, r, world)[1] - shard_range(global_batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    padded = torch.zeros((mx, *x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    padded[:x_local.shape[0]] = x_local
    out = torch.empty((mx * world, *x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx:r * mx + s] for r, s in enumerate(sizes)], dim=0)


class GradReducer:
    """Flat fp32 gradient buffer cut into buckets; a bucket is all-reduced (average) the moment its last gradient
    arrives.  Parameters are laid out in REVERSE registration order, which is the order the backward pass of the
    score net produces them, so buckets complete front to back while later kernels are still running."""

    def __init__(self, params, bucket_bytes: int = 64 << 20, group=None, comm_dtype=None):
        """comm_dtype=torch.bfloat16: a bucket is cast to bf16 before its all-reduce and back to fp32 afterwards (half
        the bytes on NVLink: 446 MB instead of 891 MB for the CelebA net).  The average then carries bf16
        rounding (about 3e-3 relative per element); parameters, Adam moments and the flat buffer stay fp32."""
        self.group = group
        self.comm_dtype = comm_dtype
        self.world, self.rank = _world(group)
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradReducer: no parameters require gradients")
        order = list(reversed(self.params))
        dev = order[0].device
        total = sum(p.numel() for p in order)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.comm = torch.zeros(total, dtype=comm_dtype, device=dev) if comm_dtype not in (None, torch.float32) else None
        self.slot, self.bucket_of, self.buckets = {}, {}, []
        off, start, count = 0, 0, 0
        for p in order:
            self.slot[p] = (off, p.numel())
            self.bucket_of[p] = len(self.buckets)
            off += p.numel()
            count += 1
            if (off - start) * 4 >= bucket_bytes:
                self.buckets.append((start, off, count))
                start, count = off, 0
        if count:
            self.buckets.append((start, off, count))
        self._use_avg = dev.type == "cuda"  # NCCL has ReduceOp.AVG; gloo sums and divides
        self.begin()

    def begin(self):
        # `loss.backward()` hands autograd VIEWS of the flat buffer; AccumulateGrad adopts them as `p.grad`.  If such a
        # gradient is still alive when the next backward starts (zero_grad(set_to_none=False), a manual p.grad.zero_(),
        # gradient accumulation over several backward passes), writing this pass's gradient into the slot would
        # overwrite `p.grad` in place and autograd would then add the slot to itself (every gradient doubled from the
        # second pass on).  Detach those gradients from the buffer first: autograd then accumulates into the private copy.
        base = self.flat.untyped_storage().data_ptr()
        for p in self.params:
            g = p.grad
            if g is not None and g.untyped_storage().data_ptr() == base:
                p.grad = g.clone()
        self._pending = [c for _, _, c in self.buckets]
        self._seen = set()
        self._works = []
        self.launched = []  # bucket ids in launch order (introspection / tests)

    def completes_bucket(self, p: torch.Tensor) -> bool:
        """Whether reporting `p` next would launch its bucket's all-reduce (producers that defer writes flush first)."""
        return self._pending[self.bucket_of[p]] == 1

    def grad_ready(self, p: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
        """Called by the backward pass with the final gradient of `p`; returns the view that will hold the average."""
        off, n = self.slot[p]
        view = self.flat[off:off + n].view(p.shape)
        if g.data_ptr() != view.data_ptr():  # producers may write straight into `grad_view(p)`
            view.copy_(g)
        if p in self._seen:
            raise RuntimeError("GradReducer: a parameter reported its gradient twice in one backward pass")
        self._seen.add(p)
        b = self.bucket_of[p]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)
        return view

    def _launch(self, b: int):
        s, e, _ = self.buckets[b]
        self.launched.append(b)
        if self.world == 1:
            return
        op = dist.ReduceOp.AVG if self._use_avg else dist.ReduceOp.SUM
        buf = self.flat[s:e]
        if self.comm is not None:
            buf = self.comm[s:e]
            buf.copy_(self.flat[s:e])  # cast on the producing stream; the collective is ordered after it
        self._works.append((dist.all_reduce(buf, op=op, group=self.group, async_op=True), s, e))

    def finish(self):
        """Flush buckets whose parameters produced no gradient (their slots are zeroed) and make the current stream
        wait for every all-reduce."""
        for p in self.params:
            if p not in self._seen:
                off, n = self.slot[p]
                self.flat[off:off + n].zero_()
                self._seen.add(p)
                b = self.bucket_of[p]
                self._pending[b] -= 1
                if self._pending[b] == 0:
                    self._launch(b)
        for work, s, e in self._works:
            work.wait()
            if self.comm is not None:
                self.flat[s:e].copy_(self.comm[s:e])
            if not self._use_avg:
                self.flat[s:e].div_(self.world)
        self._works = []

    def grad_view(self, p):
        off, n = self.slot[p]
        return self.flat[off:off + n].view(p.shape)


class DataParallelScoreNet(nn.Module):
    """Data-parallel wrapper of a score net for DSM training (train_lat_celebhq_unet_cont2.py:56-106 run on G GPUs):
    parameters are broadcast from rank 0 at construction; `loss.backward()` leaves the rank-AVERAGED gradients in
    `p.grad` (views of one flat buffer).  Same call contract as the wrapped net: `ddp(x, t)`."""

    def __init__(self, module: nn.Module, bucket_mb: float = 1024.0, process_group=None, grad_comm_dtype=None):
        """bucket_mb: size of the gradient buckets that are all-reduced while the rest of the backward pass runs.  The
        default puts every score net of this path (<= 227 M parameters) into ONE bucket, reduced right after the pass:
        measured on 8 B200s (CelebA net, 256 latents per GPU, profiles/r2_dsm_dp_ab_n8.jsonl) 24.7 ms per step against
        25.4 ms with 64 MB buckets -- the overlapped NCCL kernels take SMs from the persistent GEMM kernels, whose static
        tile schedule then waits for its slowest SM pair (capping NCCL at 8 / 4 CTAs made it 30.4 / 37.3 ms).  Pass a
        smaller value to overlap the communication of larger nets."""
        super().__init__()
        self.module = module
        self.process_group = process_group
        if _world(process_group)[0] > 1:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group else 0,
                                   group=process_group)
        self.reducer = GradReducer(module.parameters(), int(bucket_mb * (1 << 20)), process_group,
                                   comm_dtype=grad_comm_dtype)
        module._grad_sink = self.reducer  # picked up by autograd._Plan

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def state_dict(self, *args, **kwargs):  # checkpoints keep the reference's key names (no "module." prefix)
        return self.module.state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        return self.module.load_state_dict(*args, **kwargs)
