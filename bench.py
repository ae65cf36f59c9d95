#!/usr/bin/env python
"""bench.py — headline benchmark of the fused registration epoch (BASELINE.json metric:
voxel-warps/s fwd+bwd, fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1]'s volume shape, 1x1x192x192x160, NCC loss,
as a batch of PAIRS_PER_GPU independent pairs per GPU (configs[3] is the same batch, 64 pairs
sharded over 8 GPUs = 8 per GPU).  One pair (47 MB) fits the 126 MB L2, so the batch — 377 MB
streamed per step — is what makes every step read its inputs from HBM ("inputs larger than L2").
  step   = one fused epoch over the batch: ONE launch of affine_moments_kernel<3,true> that
           samples, reduces the loss moments and d(loss)/d(theta), and applies the SGD update.
  value  = voxel-warps/s with inputs resident in HBM (CUDA events, max over ranks).
  e2e    = the same metric through the public drop-in API, per pair as a reference user would:
           pinned host buffers -> Register('rigid').optim -> warp -> Register('affine').optim ->
           theta/loss read back, with the README epoch schedule scaled by --e2e-scale.
  cpu_baseline / --impl reference: the oracle port (the reference's own torch ops on the host
           cores, oracle/torch_port.py) on a bounded sample of the same workload.
Multi-GPU: pairs are independent -> batch sharding, no data-path collective; weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (192, 192, 160)
PAIRS_PER_GPU = 8
BYTES_PER_VOXEL_WARP = 8.0          # target 4 B + moving 4 B, each read once
README_EPOCHS = (500, 200)          # rigid, affine (reference README.md:59-60,70-71)
METRIC = "voxel-warps/s fwd+bwd"
UNIT = "voxel-warps/s"


def _traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_read"] + t["dram_bytes_write"]), t["source"]
    except Exception:
        return None, None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under load."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_leg(steps: int, warmup: int):
    """The reference's CPU path (oracle port: same torch ops, all host threads) on ONE pair of the
    workload; one step = one epoch (forward + backward + SGD)."""
    import torch
    from oracle import torch_port as tp
    from torchregister_b200.synth import make_pair
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mov, tgt = make_pair(SHAPE, "affine")
    p = tp.identity_params(3).clone().requires_grad_(True)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if p.grad is not None:
            p.grad = None
        warped = tp.affine_warp(p.view(1, 3, 4), mov)
        err = tp.weighted_loss(tgt, warped, (0.0, 1.0, 0.0))
        err.backward()
        with torch.no_grad():
            p -= 1e-5 * p.grad
        _ = err.item()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    vox = SHAPE[0] * SHAPE[1] * SHAPE[2]
    sec = sum(times) / len(times)
    return {"value": vox / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d epochs (after %d warm-up) of ONE %dx%dx%d pair, affine+NCC, oracle/torch_port.py "
                      "(F.affine_grid+F.grid_sample+autograd+SGD, torch %s, %d threads); %.3f s/epoch"
                      % (steps, warmup, SHAPE[0], SHAPE[1], SHAPE[2], torch.__version__, torch.get_num_threads(), sec)}, sec


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    cb, sec = cpu_reference_leg(steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference CPU path = oracle port (the pure-Python reference cannot travel to the GPU box; "
                    "its arithmetic is the same torch ops). Steps are capped at 5 epochs of one pair "
                    "(epoch time is stationary)."}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[1] shape 1x1x192x192x160, NCC loss (weight=[0,1,0]), affine epoch "
                        "(fused sample + loss + d/dtheta + SGD), batch of %d independent pairs per GPU "
                        "(configs[3] sharding: 64 pairs / 8 GPUs)" % PAIRS_PER_GPU,
            "volume": list(SHAPE), "pairs_per_gpu": PAIRS_PER_GPU, "global_pairs": PAIRS_PER_GPU * n_gpus,
            "parallelism": "batch-sharded x%d, no collective" % n_gpus,
            "l2_policy": "inputs larger than L2 (377 MB streamed per step per GPU vs 126 MB L2)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-scale", type=float, default=1.0,
                    help="fraction of the README schedule (500 rigid + 200 affine epochs) per e2e step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for the product arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import torchregister_b200 as tr
    import torchregister_b200.functional as TF
    from torchregister_b200.synth import make_pair

    W = max(3, args.warmup)
    K = max(1, args.steps)
    vox = SHAPE[0] * SHAPE[1] * SHAPE[2]

    # ---- synthetic batch resident in HBM -------------------------------------------------
    movs, tgts = [], []
    for i in range(PAIRS_PER_GPU):
        m, t = make_pair(SHAPE, "affine", seed=1234 + rank * PAIRS_PER_GPU + i, device=dev)
        movs.append(m); tgts.append(t)
    mov = torch.cat(movs).contiguous(); tgt = torch.cat(tgts).contiguous()
    del movs, tgts
    ident = torch.eye(3, 4, device=dev).reshape(1, -1)
    prob = TF.AffineProblem(mov, tgt, "affine", ident, W + K)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    prob.run(W, 1e-5, 0.0, 1.0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_load0 = time.time()
    e0.record()
    prob.run(K, 1e-5, 0.0, 1.0)           # K launches, one per step, enqueued by ONE C-ABI call
    e1.record()
    barrier()
    t_load1 = time.time()
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    losses = prob.losses[0, W:W + 3].tolist()
    value = world * PAIRS_PER_GPU * vox * K / (ms * 1e-3)

    # ---- end to end through the public API with host buffers ------------------------------
    er, ea = max(1, int(README_EPOCHS[0] * args.e2e_scale)), max(1, int(README_EPOCHS[1] * args.e2e_scale))
    host_m = mov.cpu().pin_memory(); host_t = tgt.cpu().pin_memory()
    reg0 = torch.tensor([0.02, -0.01, 0.03, 0.05, -0.05, 0.02])

    copy_stream = torch.cuda.Stream(device=dev)
    reg0_batch = reg0.repeat(PAIRS_PER_GPU, 1)

    def fetch():
        """H2D of one batch of pairs from pinned host memory on a side stream (overlaps the previous batch's epochs)."""
        with torch.cuda.stream(copy_stream):
            m = host_m.to(dev, non_blocking=True)
            t = host_t.to(dev, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream)
        return m, t, ev

    def e2e_step(cur, prefetch):
        """One batch through the public API: Register (batch extension: [N,1,D,H,W] = N independent pairs, one
        launch per epoch for all of them) rigid -> warp -> affine -> thetas to the host."""
        m, t, ev = cur
        cs = torch.cuda.current_stream(dev)
        cs.wait_event(ev)
        m.record_stream(cs); t.record_stream(cs)
        nxt = fetch() if prefetch else None
        r = tr.Register(mode="rigid", device=dev, weight=[0.0, 1.0, 0.0])
        r.optim(m, t, lr=1e-5, max_epochs=er, reg0=reg0_batch)
        m2 = r(m)
        a = tr.Register(mode="affine", device=dev, weight=[0.0, 1.0, 0.0])
        a.optim(m2, t, lr=1e-5, max_epochs=ea)
        out = torch.cat([r.theta.reshape(PAIRS_PER_GPU, -1), a.theta.reshape(PAIRS_PER_GPU, -1)], 1).to("cpu", non_blocking=True)
        return out, nxt

    _, _ = e2e_step(fetch(), False)                  # warm-up
    torch.cuda.synchronize(dev)
    barrier()
    n_e2e = 5
    t0 = time.perf_counter()
    cur = fetch()
    for i in range(n_e2e):
        res, cur = e2e_step(cur, i + 1 < n_e2e)
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * PAIRS_PER_GPU * vox * (er + ea) * n_e2e / float(dt.item())
    h2d = PAIRS_PER_GPU * 2 * vox * 4
    d2h = PAIRS_PER_GPU * 24 * 4
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    traffic, traffic_src = _traffic()
    kernel_s = ms * 1e-3 / K
    achieved = BYTES_PER_VOXEL_WARP * PAIRS_PER_GPU * vox / kernel_s / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(world),
        "iters_per_s": K / (ms * 1e-3),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "per step: pinned host -> Register('rigid').optim(%d ep) -> warp -> "
                        "Register('affine').optim(%d ep) -> thetas to host (README schedule x %.2f) on a batch of %d "
                        "pairs per Register call ([N,1,D,H,W] batch extension); %d steps; the next batch's H2D copy "
                        "overlaps the current batch's epochs (the first one does not)"
                        % (er, ea, args.e2e_scale, PAIRS_PER_GPU, n_e2e)},
        "gpu_launches": K,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel": "trb::affine3d_tma_kernel<40,20,12,4,true,false> (csrc/affine_tma.cu)",
                     "algorithmic_bytes_per_launch": BYTES_PER_VOXEL_WARP * PAIRS_PER_GPU * vox,
                     "kernel_us": kernel_s * 1e6,
                     "frac_of_nominal_8TBps": achieved / 8000.0},
        "clocks": clocks,
        "first_losses": losses,
    }

    if not args.no_extra and world == 1:
        extra = {}
        for name, shape, pairs in (("single_pair_192x192x160_L2_resident", SHAPE, 1),
                                   ("single_pair_256^3", (256, 256, 256), 1)):
            m, t = make_pair(shape, "affine", device=dev)
            pb = TF.AffineProblem(m, t, "affine", ident, 20 + 200)
            pb.run(20, 1e-5, 0.0, 1.0)
            torch.cuda.synchronize(dev)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); pb.run(200, 1e-5, 0.0, 1.0); a1.record()
            torch.cuda.synchronize(dev)
            us = a0.elapsed_time(a1) * 1e3 / 200
            v = shape[0] * shape[1] * shape[2]
            extra[name] = {"us_per_epoch": us, "iters_per_s": 1e6 / us, "voxel_warps_per_s": v / (us * 1e-6),
                           "algorithmic_GBps": 8.0 * v / (us * 1e-6) / 1e9}
            del pb, m, t
        # the reference's "criterion given -> MSE only" branch (warpings.py:38-40,125-127) on the headline batch
        pb = TF.AffineProblem(mov, tgt, "affine", ident, 20 + 100)
        pb.run(20, 1e-5, 1.0, 0.0)
        torch.cuda.synchronize(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); pb.run(100, 1e-5, 1.0, 0.0); a1.record()
        torch.cuda.synchronize(dev)
        us = a0.elapsed_time(a1) * 1e3 / 100
        extra["batch_8x192x192x160_mse_only"] = {"us_per_epoch": us, "voxel_warps_per_s": PAIRS_PER_GPU * vox / (us * 1e-6),
                                                 "algorithmic_GBps": 8.0 * PAIRS_PER_GPU * vox / (us * 1e-6) / 1e9}
        # the reference's DEFAULT loss (weights .33/.33/.33 incl. the NMI/KDE term, csrc/nmi.cu) on one pair
        one_m, one_t = mov[:1].contiguous(), tgt[:1].contiguous()
        rd = tr.Register(mode="affine", device=dev)
        rd.optim(one_m, one_t, lr=1e-5, max_epochs=3)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        rd.optim(one_m, one_t, lr=1e-5, max_epochs=30)
        torch.cuda.synchronize(dev)
        us = (time.perf_counter() - t0) / 30 * 1e6
        extra["single_pair_default_loss_mse+ncc+nmi"] = {"us_per_epoch": us, "voxel_warps_per_s": vox / (us * 1e-6),
                                                         "note": "wall clock through Register (8 + ~14 launches per epoch)"}
        line["extra"] = extra

    if not args.no_cpu_baseline and world == 1:
        cb, _ = cpu_reference_leg(2, 1)
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
