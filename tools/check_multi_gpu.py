"""Run under torchrun on >= 2 GPUs: checks the NCCL paths of the score-model hot path against single-GPU runs.
  1. batch-sharded conditional PC sampling in exact mode (2-scalar all-reduce per corrector step) + final gather
     == the unsharded batch on one GPU;
  2. data-parallel DSM step: rank-averaged gradients (bucketed all-reduce issued from inside the backward pass)
     == gradients of the same loss over the concatenated batch on one GPU;
  3. the data-parallel training step captured as ONE CUDA graph (NCCL all-reduces inside) == the eager DDP loop.
Usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tools/check_multi_gpu.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import distributed as D  # noqa: E402
from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh  # noqa: E402
from score_based_multimodal_autoencoder_b200.unet_model import Unet  # noqa: E402


def rel_max(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    kw = dict(dim=32, channels=5, dim_mults=(1, 2, 2, 2))
    model = Unet(**kw).to(dev).eval()
    sde = sh.VPSDE(1.0, 5.0, 8)
    g = torch.Generator().manual_seed(5)
    GB = 8 * world
    z = torch.randn(GB, 5, 8, 8, generator=g).to(dev)
    x0 = torch.randn(GB, 5, 8, 8, generator=g).to(dev)

    # ---- 1. sharded sampling, exact mode
    lo, hi = D.seed_shard(77, GB)
    shard = sh.cond_sampler(z[lo:hi], "0", "01234", model, sde, x_init=x0[lo:hi], global_batch=GB,
                            reduce_fn=D.corrector_allreduce())
    full_sharded = D.gather_batch(shard, GB)
    sh.manual_seed(77)
    full = sh.cond_sampler(z, "0", "01234", model, sde, x_init=x0)
    e1 = rel_max(full_sharded, full)
    # independent-shard mode (no communication in the loop) runs too and differs only through the batch means
    D.seed_shard(77, GB)
    indep = D.gather_batch(sh.cond_sampler(z[lo:hi], "0", "01234", model, sde, x_init=x0[lo:hi]), GB)
    e1b = rel_max(indep, full)

    # ---- 2. data-parallel DSM gradients
    model.train()
    u = torch.rand(GB, generator=g).to(dev)
    zz = torch.randn(GB, 5, 8, 8, generator=g).to(dev)
    ddp = D.DataParallelScoreNet(model, bucket_mb=0.25)
    loss = sh.loss_fn(z[lo:hi], ddp, sde, likelihood_weighting=False, u=u[lo:hi], z=zz[lo:hi])
    model.zero_grad(set_to_none=True)
    loss.backward()
    avg = [p.grad.detach().clone() for p in model.parameters()]
    nb = len(ddp.reducer.buckets)
    del model._grad_sink
    model.zero_grad(set_to_none=True)
    loss_full = sh.loss_fn(z, model, sde, likelihood_weighting=False, u=u, z=zz)
    loss_full.backward()
    num = sum(((a.double() - p.grad.double()) ** 2).sum() for a, p in zip(avg, model.parameters()))
    den = sum((p.grad.double() ** 2).sum() for p in model.parameters())
    e2 = (num / den).sqrt().item()
    lsum = loss.detach().clone()
    dist.all_reduce(lsum)
    e3 = abs(lsum.item() / world - loss_full.item()) / abs(loss_full.item())
    # ---- 3. the data-parallel training step as ONE CUDA graph (NCCL all-reduces captured) == the eager DDP loop
    from score_based_multimodal_autoencoder_b200.optim import FusedAdam, GraphedTrainStep
    batches = [torch.randn(8, 5, 8, 8, generator=torch.Generator().manual_seed(100 + 7 * rank + k)).to(dev) for k in range(5)]

    def fresh():
        torch.manual_seed(0)
        mm = Unet(**kw).to(dev).train()
        return mm, D.DataParallelScoreNet(mm, bucket_mb=0.25)

    m_e, d_e = fresh()
    opt = FusedAdam(m_e.parameters(), lr=5e-4)
    sh.manual_seed(31, sample_offset=rank * 8)
    eager = []
    for b in [batches[0], batches[0]] + batches:
        l = sh.loss_fn(b, d_e, sde, likelihood_weighting=False, rng="philox")
        opt.zero_grad(set_to_none=True)
        l.backward()
        opt.step()
        eager.append(l.item())
    m_g, d_g = fresh()
    sh.manual_seed(31, sample_offset=rank * 8)
    gstep = GraphedTrainStep(d_g, sde, batches[0], lr=5e-4, warmup=2)
    got = [gstep(b).item() for b in batches]
    gstep.close()  # before destroy_process_group: the graph holds NCCL kernels
    e4 = max(abs(a - b) / abs(b) for a, b in zip(got, eager[2:]))
    num = sum(((pe.double() - pg.double()) ** 2).sum() for pe, pg in zip(m_e.parameters(), m_g.parameters()))
    den = sum((pe.double() ** 2).sum() for pe in m_e.parameters())
    e4 = max(e4, (num / den).sqrt().item())
    ok = e1 < 1e-5 and e2 < 2e-3 and e3 < 1e-5 and 1e-7 < e1b < 0.2 and e4 < 5e-3
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"multi-gpu check (world {world}): sharded exact sampling rel-max {e1:.2e} (independent shards {e1b:.2e}), "
              f"DDP grads rel-L2 {e2:.2e} over {nb} buckets, loss {e3:.2e}, graphed DDP step vs eager {e4:.2e} -> "
              f"{'OK' if t.item() == 1 else 'FAIL'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
