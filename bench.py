#!/usr/bin/env python
"""Benchmark of the SBM-AE latent score-model hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload celeba_pc|poly_pc|poly_dsm]

Default workload (BASELINE.json configs[2], the config the headline metric "PC-sampler latent samples/sec" is
quoted on): CelebAMask-HQ 3-modality latent score net `Unet(dim=256, channels=3, dim_mults=(1,2,2,2,2))`,
VPSDE(0.1, 20, N=1000), conditional predictor-corrector sampling (1 of 3 modalities observed, noise_obs,
predictor -> corrector, n_steps=1, snr 0.16), 1024 latents PER GPU (batch-sharded, weak scaling: every rank samples
its own 1024 latents, no data-path collective; a reference run at N GPUs would be N such batches).
One "step" = one predictor-corrector step over the batch = 2 score-net forwards + the fused sampler kernels.
value = latent samples/s for a full N-step sample = global_batch / (N * seconds_per_step).
Synthetic latents, random-init weights (no datasets/checkpoints exist offline).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (net kwargs, latent (M, D), sde (b0, b1, N), given, all_mods, global batch)
    "celeba_pc": (dict(dim=256, channels=3, dim_mults=(1, 2, 2, 2, 2)), (3, 16), (0.1, 20.0, 1000), "0", "012", 1024),
    "poly_pc": (dict(dim=64, channels=5, dim_mults=(1, 2, 2, 2)), (5, 8), (1.0, 5.0, 100), "0", "01234", 64),
}
FWD_GFLOP_PER_SAMPLE = {"celeba_pc": 9.3496, "poly_pc": 0.1563}  # SURVEY.md 2.2 (counted on the reference modules)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_problem(workload, local_batch, rank, device):
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    kw, (M, D), (b0, b1, N), given, mods, _ = WORKLOADS[workload]
    torch.manual_seed(0)
    model = Unet(**kw).to(device).eval()
    sde = sh.VPSDE(b0, b1, N)
    g = torch.Generator().manual_seed(1234 + rank)
    z_host = torch.randn(local_batch, M, D, D, generator=g).pin_memory()
    x_host = torch.randn(local_batch, M, D, D, generator=g).pin_memory()
    return model, sde, z_host, x_host, given, mods


def run_ours(args):
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
    import torch.distributed as dist
    from score_based_multimodal_autoencoder_b200 import _lib as L
    from score_based_multimodal_autoencoder_b200 import ops
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    saved_stdout = None
    if world > 1:
        # NCCL / c10d print a version banner on stdout when the first communicator comes up: keep stdout for the one
        # JSON line by pointing fd 1 at stderr until the result is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=device)
    workload = args.workload
    kw, (M, D), (b0, b1, N), given, mods, local_batch = WORKLOADS[workload]
    if args.batch:
        local_batch = args.batch
    # weak scaling: every rank samples its own `local_batch` latents (independent shards, no data-path collective)
    global_batch = local_batch * world
    model, sde, z_host, x_host, given, mods = build_problem(workload, local_batch, rank, device)
    sh.manual_seed(20240607, sample_offset=rank * local_batch)
    mask = sh._obs_mask_from(given, mods)
    z_obs = z_host.to(device)
    x0 = x_host.to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timed region: K consecutive PC steps of the conditional sampler
    def run_steps(n, x, use_graph):
        return sh.pc_sampler(x, model, sde, z_obs=z_obs, obs_mask=mask, num_steps=n, use_graph=use_graph,
                             return_state=True)[1]

    state = run_steps(max(args.warmup, 3), x0, False)  # warm-up: packs weights, sizes the allocator
    barrier()
    stepper = _GraphStepper(sh, model, sde, z_obs, mask, state) if args.graph else None
    if stepper is not None:
        for _ in range(3):
            stepper.step()
    barrier()
    launches0 = L.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record()
        if stepper is not None:
            for _ in range(args.steps):
                stepper.step()
        else:
            state = run_steps(args.steps, state, False)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = L.launch_count() - launches0
    if stepper is not None:
        launches = stepper.launches_per_step * args.steps
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = t.item() / args.steps
    value = global_batch / (N * ms_per_step * 1e-3)

    # ---------------- end to end through the public API with HOST buffers (H2D + step + D2H every step)
    out_host = torch.empty_like(x_host).pin_memory()
    def e2e_step():
        zo = z_host.to(device, non_blocking=True)
        xi = x_host.to(device, non_blocking=True)
        out = sh.cond_sampler(zo, given, mods, model, sde, x_init=xi, num_steps=1)
        out_host.copy_(out, non_blocking=True)
    for _ in range(2):
        e2e_step()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        e2e_step()
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / args.steps
    e2e = {"value": global_batch / (N * e2e_ms * 1e-3), "unit": "samples/s",
           "h2d_bytes_per_step": 2 * z_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
           "ms_per_step": e2e_ms}

    # ---------------- roofline of the dominant kernel family (tcgen05 implicit-GEMM convolutions), timed live
    roof = conv_roofline(model, sde, x0, ops, L) if rank == 0 else None
    samp = sampler_kernel_roofline(sh, sde, device) if rank == 0 else None
    dsm = None
    if not args.no_dsm:
        del stepper
        torch.cuda.empty_cache()
        dsm = dsm_train_bench(device, max(args.steps, 5), args.warmup, "poly" if world == 1 else "celeba", world, rank,
                              bool(args.graph))

    line = None
    if rank == 0:
        pk, pk_src = peaks()
        fwd_flops = FWD_GFLOP_PER_SAMPLE[workload] * 1e9 * local_batch
        line = {
            "metric": "pc_sampler_latent_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic latents N(0,1), random-init weights (torch.manual_seed(0))",
            "config": {"workload": f"{workload}: Unet{tuple(kw.values())} cond. PC sampling given={given!r} of {mods!r}, "
                                   f"VPSDE({b0},{b1},N={N}), n_steps=1, snr=0.16, predictor->corrector",
                       "global_batch": global_batch, "per_gpu_batch": local_batch, "latent": [M, D, D],
                       "sde_steps_per_sample": N, "step": "1 PC step = 2 score-net forwards + fused sampler kernels",
                       "parallelism": f"batch-sharded x{world}", "cuda_graph": bool(args.graph),
                       "l2": "activations and weights per forward exceed the 126 MB L2 (no flush needed)"
                             if workload == "celeba_pc" else "L2-resident working set (latency-bound config)"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary(),
            "net_tflops_per_gpu": 2 * fwd_flops / (ms_per_step * 1e-3) / 1e12,
            "roofline": roof and {**roof, "peak": pk["bf16_tflops_sustained"],
                                  "frac": roof["achieved"] / pk["bf16_tflops_sustained"], "peak_source": pk_src},
            "roofline_sampler_kernels": samp and {**samp, "peak": pk["hbm_gbs"], "frac": samp["achieved"] / pk["hbm_gbs"],
                                                  "peak_source": pk_src},
        }
        if dsm is not None:
            line["dsm_train"] = dsm
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(workload, budget_s=20.0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


class _GraphStepper:
    """One predictor-corrector step captured as a CUDA graph (per-step state advanced on the device)."""

    def __init__(self, sh, model, sde, z_obs, mask, x):
        import ctypes as C
        from score_based_multimodal_autoencoder_b200 import _lib as L
        self.sh, self.L, self.C = sh, L, C
        dev = x.device
        B = x.shape[0]
        self.ts = torch.linspace(sde.T, 1e-3, sde.N, device=dev)
        self.step_dev = torch.full((1,), 3, dtype=torch.int32, device=dev)
        self.draw_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.t_next = torch.zeros(1, dtype=torch.float32, device=dev)
        self.t_vec = torch.empty(B, dtype=torch.float32, device=dev)
        self.x = x.clone()
        self.acc = torch.zeros(3, dtype=torch.float64, device=dev)
        self.model, self.sde, self.z_obs, self.mask = model, sde, z_obs, mask
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = L.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self._body()
        self.launches_per_step = L.launch_count() - n0

    def _tick(self, advance):
        L, C = self.L, self.C
        L.check(L.lib().sbm_sampler_tick(L.ptr(self.ts), C.c_int32(self.ts.numel()), L.ptr(self.step_dev),
                                         L.ptr(self.draw_dev), L.ptr(self.t_vec), C.c_int32(self.x.shape[0]),
                                         L.ptr(self.t_next), C.c_int32(advance), C.c_uint64(2), L.stream_ptr()),
                "sbm_sampler_tick")

    def _body(self):
        sh = self.sh
        self._tick(0)
        im = sh._impute_struct(self.z_obs, self.mask, True, 0.0, self.t_next)
        rng = sh._RngState()
        score = self.model(self.x, self.t_vec)
        x1, _ = sh._predictor_kernel(self.sde, self.x, score, self.t_vec, rng=rng.next(self.draw_dev), want_mean=False)
        grad = self.model(x1, self.t_vec)
        sh._corrector_kernels(self.sde, x1, grad, self.t_vec, 0.16, rng=rng.next(self.draw_dev), impute=im,
                              want_mean=False, acc=self.acc, out=self.x)
        self._tick(1)

    def step(self):
        self.graph.replay()


def dsm_train_bench(device, steps, warmup, which="poly", world=1, rank=0, use_graph=True):
    """DSM training step.  which="poly": BASELINE configs[1] (PolyMNIST latent score UNet, batch 256, 1 GPU);
    which="celeba": configs[3] (CelebAMask-HQ latent UNet, data parallel, 256 latents per GPU = weak scaling, bucketed
    NCCL gradient all-reduce overlapped with the hand-written backward).  bf16 GEMM operands / fp32 master weights.
    One step = loss_fn (fused perturb, net forward, fused loss) + backward + FusedAdam.  Collective: call on all ranks."""
    import torch.distributed as dist
    from score_based_multimodal_autoencoder_b200 import _lib as L
    from score_based_multimodal_autoencoder_b200 import sde_helper2 as sh
    from score_based_multimodal_autoencoder_b200.distributed import DataParallelScoreNet
    from score_based_multimodal_autoencoder_b200.optim import FusedAdam
    from score_based_multimodal_autoencoder_b200.unet_model import Unet
    if which == "poly":
        kw, shape, sde, lr, fwd_gf = dict(dim=64, channels=5, dim_mults=(1, 2, 2, 2)), (256, 5, 8, 8), sh.VPSDE(1.0, 5.0, 100), 5e-4, 0.1563
    else:
        kw, shape, sde, lr, fwd_gf = dict(dim=256, channels=3, dim_mults=(1, 2, 2, 2, 2)), (256, 3, 16, 16), sh.VPSDE(0.1, 20.0, 1000), 5e-5, 9.3496
    torch.manual_seed(0)
    model = Unet(**kw).to(device).train()
    net = DataParallelScoreNet(model, bucket_mb=64.0) if world > 1 else model
    opt = FusedAdam(model.parameters(), lr=lr)
    sh.manual_seed(777, sample_offset=rank * shape[0])
    z_host = torch.randn(*shape, generator=torch.Generator().manual_seed(1234 + rank)).pin_memory()
    z = z_host.to(device)

    def eager_step(batch):
        loss = sh.loss_fn(batch, net, sde, reduce_mean=True, likelihood_weighting=False, eps=1e-5, rng="philox")
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    graphed, step = None, eager_step
    if use_graph:
        # the whole step (loss_fn + backward [+ bucketed NCCL all-reduces] + FusedAdam) replayed as ONE CUDA graph;
        # per-step state (Philox draw id, Adam step count) lives on the device
        from score_based_multimodal_autoencoder_b200.optim import GraphedTrainStep
        try:
            graphed = GraphedTrainStep(net, sde, z, lr=lr, warmup=3)
            step = graphed
        except Exception as exc:  # e.g. an NCCL build that cannot be captured: fall back to the eager loop
            if rank == 0:
                print(f"[bench] CUDA-graph capture of the training step failed ({type(exc).__name__}: {exc}); eager loop",
                      file=sys.stderr)
            graphed, step = None, eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps

    for _ in range(max(warmup, 3)):
        loss = step(z)
    barrier()
    n0 = L.launch_count()
    ms = timed(lambda: step(z))
    launches = graphed.launches_per_step if graphed is not None else (L.launch_count() - n0) // steps
    launches = int(launches)
    # end to end: H2D of the latent batch and D2H of the loss every step (the reference does loss.item() per step)
    ms_e2e = timed(lambda: step(z_host.to(device, non_blocking=True)).item())
    loss_val = float(step(z).item())
    used_graph = graphed is not None
    # a captured graph that contains NCCL kernels must be destroyed BEFORE the process group (destroying the
    # communicator first dead-locks at exit)
    step = graphed = None
    torch.cuda.synchronize()
    gb = shape[0] * world
    return {"metric": "dsm_train_steps_per_sec", "value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms,
            "n_gpus": world, "scaling": "weak", "global_batch": gb, "latents_per_sec": gb * 1e3 / ms,
            "e2e": {"value": 1e3 / ms_e2e, "unit": "steps/s", "h2d_bytes_per_step": z_host.numel() * 4 * world,
                    "d2h_bytes_per_step": 4 * world},
            "gpu_launches_per_step": launches, "loss": loss_val, "cuda_graph": used_graph,
            "model_tflops_per_gpu": 3 * fwd_gf * 1e9 * shape[0] / (ms * 1e-3) / 1e12,
            "grad_allreduce": None if world == 1 else {"backend": "nccl", "bucket_mb": 64,
                                                       "bytes_per_step": sum(p.numel() for p in model.parameters()) * 4,
                                                       "overlap": "buckets launched from inside the backward pass"},
            "config": {"workload": f"{which}_dsm: Unet{tuple(kw.values())} DSM training, batch {shape[0]} per GPU x {world}, "
                                   f"latent {list(shape[1:])}, Adam lr {lr}, bf16 GEMM operands / fp32 master weights, "
                                   f"loss and statistics fp32/fp64"}}


def conv_roofline(model, sde, x0, ops, L):
    """Time every tcgen05 conv launch of ONE score-net forward with CUDA events on the launch stream.  `roofline` is the
    dominant kernel (the persistent CTA-pair kernel with 256-wide N tiles): algorithmic FLOPs (valid taps only, true
    channel counts) of its launches / their summed durations; `all_convs` is the same over every conv launch."""
    rec = []
    orig = ops.conv_igemm

    def timed(x, wpk, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(x, wpk, **kw)
        e1.record()
        variant = L.lib().sbm_conv_last_variant()
        b, h, w, _ = x.shape
        kind, kh, kwd = kw["kind"], kw["kh"], kw["kw"]
        dense = None
        if kind == L.CONV_S1:
            taps = sum(1 for i in range(kh) for j in range(kwd) if abs(i - kh // 2) < h and abs(j - kwd // 2) < w)
            pix = b * h * w
            if variant & (1 << 18):  # pixel-major tiling: taps on zero padding are skipped per output pixel, so they
                # leave the numerator too; `dense` keeps the count a dense 'same' convolution would do
                vh = sum(1 for i in range(h) for a in range(kh) if 0 <= i + a - kh // 2 < h)
                vw = sum(1 for j in range(w) for a in range(kwd) if 0 <= j + a - kwd // 2 < w)
                dense = 2.0 * pix * kw["cin"] * kw["cout"] * taps
                rec.append((e0, e1, 2.0 * b * vh * vw * kw["cin"] * kw["cout"], variant, dense))
                return out
        elif kind == L.CONV_S2:
            taps, pix = kh * kwd, b * (h // 2) * (w // 2)
            if h == 2:
                taps = 4
        else:
            taps, pix = (4 if h > 1 else 1), b * h * w * 4
        fl = 2.0 * pix * kw["cin"] * kw["cout"] * taps
        rec.append((e0, e1, fl, variant, fl))
        return out

    t = torch.full((x0.shape[0],), 0.5, device=x0.device)
    with torch.no_grad():
        model(x0, t)
        torch.cuda.synchronize()
        ops.conv_igemm = timed
        try:
            model(x0, t)
        finally:
            ops.conv_igemm = orig
    torch.cuda.synchronize()

    def agg(rows):
        ms = sum(r[0].elapsed_time(r[1]) for r in rows)
        fl = sum(r[2] for r in rows)
        return ms, fl

    # same kernel template with or without pixel-major tiling (bit 18)
    dom = [r for r in rec if (r[3] & ~(1 << 18)) == (256 | (1 << 16) | (1 << 17))] or rec
    ms_d, fl_d = agg(dom)
    ms_a, fl_a = agg(rec)
    dense_a = sum(r[4] for r in rec)
    top = max(rec, key=lambda r: r[2])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
        traffic = json.load(open(tpath)).get("conv_igemm_pair_kernel<256,5,staged>", {}).get("avg_dram_bytes_per_launch")
    return {"bound": "tensor", "kernel": "conv_igemm_pair_kernel<256,5,staged> (tcgen05 cta_group::2 implicit GEMM)",
            "achieved": fl_d / (ms_d * 1e-3) / 1e12, "unit": "TFLOP/s", "launches": len(dom),
            "algorithmic_gflop_per_launch": fl_d / 1e9 / len(dom), "avg_launch_us": ms_d * 1e3 / len(dom),
            "share_of_forward_conv_time": ms_d / ms_a,
            "largest_launch_tflops": top[2] / (top[0].elapsed_time(top[1]) * 1e-3) / 1e12, "traffic": traffic,
            "pixel_major_launches": sum(1 for r in dom if r[3] & (1 << 18)),
            "all_convs": {"achieved": fl_a / (ms_a * 1e-3) / 1e12, "launches": len(rec),
                          "algorithmic_gflop_per_forward": fl_a / 1e9, "ms_per_forward": ms_a,
                          "dense_equivalent_tflops": dense_a / (ms_a * 1e-3) / 1e12,
                          "note": "algorithmic = multiply-adds on real pixels (zero-padding taps the pixel-major "
                                  "tiling skips are not counted); dense_equivalent counts them as a dense conv does"}}


def sampler_kernel_roofline(sh, sde, device, batch=65536):
    """HBM roofline of the fused sampler-step kernels on the large-batch sweep point (64k Poly latents):
    28 B per latent element per PC step (predictor 12 + norms 4 + update 12), SURVEY.md 8(d)."""
    M, D = 5, 8
    x = torch.randn(batch, M, D, D, device=device)
    s = torch.randn_like(x)
    t = torch.full((batch,), 0.5, device=device)
    rng = sh._RngState()
    acc = torch.zeros(3, dtype=torch.float64, device=device)
    out = torch.empty_like(x)

    def pc_kernels():
        x1, _ = sh._predictor_kernel(sde, x, s, t, rng=rng.next(), want_mean=False, out=out)
        sh._corrector_kernels(sde, x1, s, t, 0.16, rng=rng.next(), want_mean=False, acc=acc, out=out)

    for _ in range(3):
        pc_kernels()
    # the kernels of 4 PC steps as one CUDA graph (how pc_sampler(use_graph=True) runs them): the timed region holds
    # kernel time, not the CPU launch gaps of an eager loop (~3 us per launch against ~40 us kernels)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timed_as = "cuda graph replay"
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            pc_kernels()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        per_graph = 4
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            for _ in range(per_graph):
                pc_kernels()
        graph.replay()
        torch.cuda.synchronize()
        reps = 5
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * per_graph)
        del graph
    except Exception as exc:  # noqa: BLE001 - measurement aid only: fall back to the eager launch loop
        print(f"[bench] sampler-kernel graph capture failed ({exc}); timing the eager loop", file=sys.stderr)
        timed_as = "eager launch loop"
        torch.cuda.synchronize()
        reps = 20
        e0.record()
        for _ in range(reps):
            pc_kernels()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    nbytes = 28.0 * x.numel()
    return {"bound": "hbm", "kernel": "predictor + corrector_norms + corrector_update", "achieved": nbytes / (ms * 1e-3) / 1e9,
            "unit": "GB/s", "us_per_pc_step": ms * 1e3, "batch": batch, "algorithmic_bytes_per_step": nbytes,
            "timed_as": timed_as,
            "note": "3 launches per PC step; the update kernel re-zeroes the norm accumulator itself"}


# --------------------------------------------------------------------------------------- CPU arms
def _oracle_problem(workload, batch):
    from oracle import sde_oracle as so
    from oracle import unet_oracle as uo
    from oracle.det_weights import fill_state_dict
    kw, (M, D), (b0, b1, N), given, mods, _ = WORKLOADS[workload]
    sd = fill_state_dict(uo.unet_param_shapes(kw["dim"], kw["channels"], kw["dim_mults"]))
    spec = so.SdeSpec("vp", b0, b1, N)
    score_fn = lambda x, t: uo.unet_forward(sd, x, t, dim=kw["dim"], dim_mults=kw["dim_mults"])
    g = torch.Generator().manual_seed(1234)
    z = torch.randn(batch, M, D, D, generator=g)
    mask = [m in given for m in mods]
    return so, spec, score_fn, z, mask, N


def _oracle_pc_steps(so, spec, score_fn, z, mask, nsteps, g):
    npred = torch.randn(nsteps, *z.shape, generator=g)
    ncorr = torch.randn(nsteps, 1, *z.shape, generator=g)
    with torch.no_grad():
        so.pc_sampler(spec, score_fn, z, npred, ncorr, z_obs=z, obs_mask=mask, num_steps=nsteps)


def cpu_baseline(workload, budget_s=20.0):
    """The oracle port (CPU fp32 restatement of the reference path) on a bounded sample of the same workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = 4 if workload == "celeba_pc" else 64
    so, spec, score_fn, z, mask, N = _oracle_problem(workload, batch)
    g = torch.Generator().manual_seed(1)
    t0 = time.time()
    _oracle_pc_steps(so, spec, score_fn, z, mask, 1, g)  # warm-up
    warm = time.time() - t0
    n = max(1, min(8, int(budget_s / max(warm, 1e-3)) - 1))
    t0 = time.time()
    _oracle_pc_steps(so, spec, score_fn, z, mask, n, g)
    dt = (time.time() - t0) / n
    return {"value": batch / (N * dt), "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{n} PC steps of batch {batch} (oracle, torch CPU fp32, {cores} threads), {dt * 1e3:.0f} ms/step",
            "ms_per_step_at_sample_batch": dt * 1e3}


def run_reference(args):
    """--impl reference: the reference's own CPU path for this workload (the oracle port, since the reference is
    Python/torch and /root/reference does not exist on the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    workload = args.workload
    kw, (M, D), (b0, b1, N), given, mods, global_batch = WORKLOADS[workload]
    global_batch *= int(os.environ.get("WORLD_SIZE", "1"))
    batch = 4 if workload == "celeba_pc" else 64
    so, spec, score_fn, z, mask, N = _oracle_problem(workload, batch)
    g = torch.Generator().manual_seed(1)
    t0 = time.time()
    _oracle_pc_steps(so, spec, score_fn, z, mask, 1, g)
    probe = time.time() - t0
    total = args.steps + args.warmup
    if probe * total > 150 and batch > 1:  # keep the whole arm within a few minutes
        batch = max(1, int(batch * 150 / (probe * total)))
        so, spec, score_fn, z, mask, N = _oracle_problem(workload, batch)
    _oracle_pc_steps(so, spec, score_fn, z, mask, max(args.warmup, 1), g)
    t0 = time.time()
    _oracle_pc_steps(so, spec, score_fn, z, mask, args.steps, g)
    dt = (time.time() - t0) / args.steps
    value = batch / (N * dt)
    sample = f"{args.steps} PC steps of batch {batch} (oracle port of the reference path, torch CPU fp32)"
    print(json.dumps({
        "impl": "reference", "metric": "pc_sampler_latent_samples_per_sec", "value": value, "unit": "samples/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic latents N(0,1), random-init weights (torch.manual_seed(0))",
        "config": {"workload": f"{workload}: same net / SDE / sampler settings as the B200 arm", "global_batch": global_batch,
                   "sample_batch": batch, "sde_steps_per_sample": N},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="celeba_pc", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--graph", type=int, default=1, help="replay one captured CUDA graph per PC step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dsm", action="store_true", help="skip the secondary DSM-training measurement")
    ap.add_argument("--dsm-only", default="", help="poly|celeba: run only the DSM training measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.dsm_only:
        import torch.distributed as dist
        world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=device)
        res = dsm_train_bench(device, args.steps, args.warmup, args.dsm_only, world, rank, bool(args.graph))
        if rank == 0:
            print(json.dumps(res), flush=True)
        if world > 1:
            dist.destroy_process_group()
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
