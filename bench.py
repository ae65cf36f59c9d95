#!/usr/bin/env python
"""bench.py — headline benchmark of the fused registration epoch (BASELINE.json metric:
voxel-warps/s fwd+bwd, fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1]'s volume shape, 1x1x192x192x160, NCC loss,
as a batch of PAIRS_PER_GPU independent pairs per GPU (configs[3] is the same batch, 64 pairs
sharded over 8 GPUs = 8 per GPU).  One pair (47 MB) fits the 126 MB L2, so the batch — 377 MB
streamed per step — is what makes every step read its inputs from HBM ("inputs larger than L2").
  step   = one fused epoch over the batch (sample, loss moments, d loss / d theta, SGD update for
           every pair).  The K timed steps are enqueued by ONE C-ABI call and run inside the
           persistent kernel affine3d_persist_kernel (csrc/affine_persist.cu): ceil(K/249) cooperative
           launches + one target_sums_kernel, no per-epoch launch.
  value  = voxel-warps/s with inputs resident in HBM (CUDA events, max over ranks).
  e2e    = the same metric through the public drop-in API: pinned host buffers -> Register('rigid')
           .optim -> warp -> Register('affine').optim -> theta read back, README epoch schedule.
  cpu_baseline / --impl reference: the UNMODIFIED reference (baseline/_ref, imported through
           oracle/ref_shim.py) on the host cores on a bounded sample of the same workload; the oracle
           port (oracle/torch_port.py) only if the reference copy is missing.
Multi-GPU: pairs are independent -> batch sharding, no data-path collective; weak scaling.  For N > 1 the
line also carries `extra.sharded_512` (BASELINE configs[4]: one 512^3 volume in z-slabs, peer-memory and NCCL
forms, speed-up against the 1-GPU time measured in the same job), `extra.configs3_strong` (64 pairs in total,
rigid -> affine -> flow) and, at N == 2, the result of the multi-GPU parity checks (tests/mgpu_check.py).
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
CHUNK_EPOCHS = 249                  # epochs per persistent launch (accumulator region of the workspace)
# Second roofline of the epoch kernel (profiles/r02_microbench_ffma2_operands.txt): the packed-fp32
# stream of one voxel-pair step costs ~158 cycles per sub-partition at the measured operand-delivery rates
# (register-file reads, not DRAM); a 32x16x16 tile is 32 such steps per sub-partition.
STEP_CYCLES_BOUND = 158.0
SM_MHZ = 1965.0


def _traffic():
    """DRAM bytes per epoch of the dominant kernel from the committed ncu capture (profiles/)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            return float(t["dram_bytes_read"] + t["dram_bytes_write"]), t["source"]
        except Exception:
            continue
    return None, None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def issue_bound_us(shape, pairs, sms=148):
    tiles = pairs * ((shape[2] + 31) // 32) * ((shape[1] + 15) // 16) * ((shape[0] + 15) // 16)
    per_cta = -(-tiles // sms)
    return per_cta * 32 * STEP_CYCLES_BOUND / SM_MHZ


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


# ------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_leg(epochs: int):
    """The reference's CPU path on ONE pair of the workload (affine, NCC through weight=[0,1,0]); one step = one epoch
    (forward + backward + SGD).  kind "reference": the UNMODIFIED reference's affine_register from baseline/_ref;
    its per-epoch time is (T(1+epochs) - T(1)) / epochs, which removes its set-up (the inert 1.18M->64 Linear and
    random.sample).  Harness-side deviation: NMILoss — evaluated by the reference even at weight 0 — is replaced by a
    zero stub, because its 3-D KDE materialises tens of GB (utils.py:24-30,242-247); everything else is stock.
    kind "port": oracle/torch_port.py (same torch ops) when no reference copy is present."""
    import torch
    from torchregister_b200.synth import make_pair
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mov, tgt = make_pair(SHAPE, "affine")
    vox = SHAPE[0] * SHAPE[1] * SHAPE[2]
    kind, sec, how = "port", None, ""
    try:
        import contextlib
        import io
        import random
        import torch.nn as nn
        from oracle import ref_shim
        if not ref_shim.available():
            raise RuntimeError("no reference copy")
        _, rw, _ = ref_shim.load()

        class _ZeroNMI(nn.Module):
            def forward(self, y, yp):
                return yp.sum() * 0

        real_nmi = rw.NMILoss
        rw.NMILoss = _ZeroNMI

        def run(n):
            random.seed(0)
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                rw.affine_register(mov, tgt, lr=1e-5, epochs=n, per=0.1, device="cpu", debug=False,
                                   weights=[0.0, 1.0, 0.0], grad_edges=False)
            return time.perf_counter() - t0
        try:
            run(1)                                       # warm-up (thread pools, allocator)
            t1 = run(1)
            t2 = run(1 + epochs)
        finally:
            rw.NMILoss = real_nmi
        sec = max((t2 - t1) / epochs, 1e-9)
        kind = "reference"
        how = ("UNMODIFIED reference affine_register(weights=[0,1,0]) from %s, %d epochs by difference T(%d)-T(1) "
               "(set-up %.2f s excluded), NMI term (weight 0) stubbed" % (os.path.relpath(ref_shim.source(), ROOT), epochs, 1 + epochs, t1))
    except Exception as e:                               # reference copy missing: the port
        from oracle import torch_port as tp
        p = tp.identity_params(3).clone().requires_grad_(True)
        times = []
        for it in range(1 + epochs):
            t0 = time.perf_counter()
            if p.grad is not None:
                p.grad = None
            warped = tp.affine_warp(p.view(1, 3, 4), mov)
            err = tp.weighted_loss(tgt, warped, (0.0, 1.0, 0.0))
            err.backward()
            with torch.no_grad():
                p -= 1e-5 * p.grad
            _ = err.item()
            if it >= 1:
                times.append(time.perf_counter() - t0)
        sec = sum(times) / len(times)
        how = "oracle/torch_port.py (F.affine_grid+F.grid_sample+autograd+SGD), %d epochs after 1 warm-up; reference unavailable: %r" % (epochs, e)
    return {"value": vox / sec, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "ONE %dx%dx%d pair, affine+NCC; %s; torch %s, %d threads; %.3f s/epoch"
                      % (SHAPE[0], SHAPE[1], SHAPE[2], how, torch.__version__, torch.get_num_threads(), sec)}, sec


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    cb, sec = cpu_reference_leg(steps)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference arm = the reference's own CPU implementation of the path on the host cores (rank 0 only). "
                    "Steps are capped at 5 epochs of one pair (epoch time is stationary)."}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[1] shape 1x1x192x192x160, NCC loss (weight=[0,1,0]), affine epoch "
                        "(fused sample + loss + d/dtheta + SGD), batch of %d independent pairs per GPU "
                        "(configs[3] sharding: 64 pairs / 8 GPUs)" % PAIRS_PER_GPU,
            "volume": list(SHAPE), "pairs_per_gpu": PAIRS_PER_GPU, "global_pairs": PAIRS_PER_GPU * n_gpus,
            "parallelism": "batch-sharded x%d, no collective" % n_gpus,
            "l2_policy": "inputs larger than L2 (377 MB streamed per step per GPU vs 126 MB L2)"}


# ------------------------------------------------------------------------------------------------------------
# helpers for the product arm
# ------------------------------------------------------------------------------------------------------------
def time_epochs(TF, torch, dev, mov, tgt, mode, p0, epochs, w, warm=20, repeats=1, optimiser="sgd"):
    prob = TF.AffineProblem(mov, tgt, mode, p0, warm + repeats * epochs)
    prob.run(warm, 1e-5, w[0], w[1], optimiser=optimiser)
    torch.cuda.synchronize(dev)
    best = None
    for _ in range(repeats):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); prob.run(epochs, 1e-5, w[0], w[1], optimiser=optimiser); a1.record()
        torch.cuda.synchronize(dev)
        us = a0.elapsed_time(a1) * 1e3 / epochs
        best = us if best is None else min(best, us)
    return best, prob


def roofline_entry(us, shape, pairs, peak, kernel):
    vox = shape[0] * shape[1] * shape[2] * pairs
    ach = BYTES_PER_VOXEL_WARP * vox / (us * 1e-6) / 1e9
    bound = issue_bound_us(shape, pairs)
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "kernel": kernel, "kernel_us_per_epoch": us, "iters_per_s": 1e6 / us,
            "algorithmic_bytes_per_epoch": BYTES_PER_VOXEL_WARP * vox, "frac_of_nominal_8TBps": ach / 8000.0,
            "issue_bound_us_per_epoch": bound, "issue_bound_frac": bound / us,
            "issue_bound_note": "packed-fp32 operand-delivery bound of the epoch kernel (158 cycles per voxel-pair step and "
                                "sub-partition from tools/microbench2.cu rates): the kernel cannot go below this however "
                                "fast HBM is; issue_bound_frac = that bound / measured"}


def sharded_512(torch, dist, dev, rank, world, epochs=40):
    """BASELINE configs[4]: ONE 512^3 volume in z-slabs over the ranks; affine NCC and direct flow, fused peer-memory
    form and NCCL form, and the same volume on ONE GPU (rank 0, same job) for the speed-up."""
    import torchregister_b200.functional as TF
    from torchregister_b200.parallel import ShardedAffine, ShardedDirectFlow
    from torchregister_b200.synth import make_pair
    shape = (512, 512, 512)
    mov, tgt = make_pair(shape, "flow", device=dev)
    vox = shape[0] * shape[1] * shape[2]
    ident = torch.eye(3, 4, device=dev).reshape(1, -1)
    out = {"volume": list(shape), "epochs": epochs}

    def timed(fn):
        fn(3)
        torch.cuda.synchronize(dev)
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(epochs); b.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a.elapsed_time(b) / epochs], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # one GPU, same job (rank 0 works, the others wait at the barrier)
    single = {}
    if rank == 0:
        us, _ = time_epochs(TF, torch, dev, mov, tgt, "affine", ident, epochs, (0.0, 1.0), warm=5)
        single["affine_ncc"] = us * 1e-3
        for opt in ("sgd", "adam"):
            sd = TF.DirectFlowProblem(mov, tgt, 3 + epochs, optimiser=opt)
            sd.run(3, 0.05, 0.5, 0.5, 2.0)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); sd.run(epochs, 0.05, 0.5, 0.5, 2.0); b.record()
            torch.cuda.synchronize(dev)
            single["direct_flow_%s" % opt] = a.elapsed_time(b) / epochs
            del sd
            torch.cuda.empty_cache()
    dist.barrier()
    s = torch.tensor([single.get("affine_ncc", 0.0), single.get("direct_flow_sgd", 0.0), single.get("direct_flow_adam", 0.0)],
                     device=dev, dtype=torch.float64)
    dist.broadcast(s, 0)
    single = {"affine_ncc": float(s[0]), "direct_flow_sgd": float(s[1]), "direct_flow_adam": float(s[2])}
    out["one_gpu_ms_per_epoch"] = single
    for name, peer in (("affine_ncc_nccl", False), ("affine_ncc_peer", None)):
        sa = ShardedAffine(mov, tgt, "affine", ident, 100000, peer=peer)
        ms = timed(lambda n: sa.run(n, 1e-5, 0.0, 1.0, align=False))
        sa.check()
        out[name] = {"ms_per_epoch": ms, "speedup_vs_1gpu": single["affine_ncc"] / ms, "voxel_warps_per_s": vox / (ms * 1e-3),
                     "collective": "peer-mailbox (in-kernel, persistent kernel)" if sa.mailbox is not None else "nccl all-reduce of 41 fp64 between two kernels"}
        del sa
    for opt in ("sgd", "adam"):
        for tag, peer in (("nccl", False), ("peer", None)):
            sd = ShardedDirectFlow(mov, tgt, 100000, optimiser=opt, peer=peer)
            ms = timed(lambda n: sd.run(n, 0.05, 0.5, 0.5, 2.0))
            sd.check()
            out["direct_flow_%s_%s" % (opt, tag)] = {
                "ms_per_epoch": ms, "speedup_vs_1gpu": single["direct_flow_%s" % opt] / ms, "voxel_warps_per_s": vox / (ms * 1e-3),
                "collective": "peer-mailbox + halo slices read in place over NVLink" if sd.mailbox is not None else "nccl (batch_isend_irecv halo + all-reduce of 6 fp64)"}
            del sd
            torch.cuda.empty_cache()
    return out


def configs3_strong(torch, dist, dev, rank, world, total_pairs=64, scale=0.2, flow_epochs=20):
    """BASELINE configs[3] as written: 64 pairs IN TOTAL, rigid -> affine -> flow per pair, batch-sharded (strong scaling).
    The flow stage is the direct per-voxel flow (SGD, MSE+NCC+smoothness); the U-Net parametrisation is replicas-only."""
    import torchregister_b200 as tr
    import torchregister_b200.functional as TF
    from torchregister_b200.parallel import shard_pairs
    from torchregister_b200.synth import make_pair
    a, b = shard_pairs(total_pairs, world, rank)
    n = b - a
    er, ea = max(1, int(README_EPOCHS[0] * scale)), max(1, int(README_EPOCHS[1] * scale))
    movs, tgts = [], []
    for i in range(a, b):
        m, t = make_pair(SHAPE, "affine", seed=5000 + i, device=dev)
        movs.append(m); tgts.append(t)
    mov, tgt = torch.cat(movs).contiguous(), torch.cat(tgts).contiguous()
    del movs, tgts
    reg0 = torch.tensor([0.02, -0.01, 0.03, 0.05, -0.05, 0.02], device=dev).repeat(n, 1)

    def pipeline():
        r = tr.Register(mode="rigid", device=dev, weight=[0.0, 1.0, 0.0])
        r.optim(mov, tgt, lr=1e-5, max_epochs=er, reg0=reg0)
        m2 = r(mov)
        af = tr.Register(mode="affine", device=dev, weight=[0.0, 1.0, 0.0])
        af.optim(m2, tgt, lr=1e-5, max_epochs=ea)
        m3 = af(m2)
        for i in range(n):
            fp = TF.DirectFlowProblem(m3[i:i + 1], tgt[i:i + 1], flow_epochs, optimiser="sgd")
            fp.run(flow_epochs, 0.05, 0.5, 0.5, 2.0)
        return af.theta
    pipeline()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pipeline(); e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    vox = SHAPE[0] * SHAPE[1] * SHAPE[2]
    return {"total_pairs": total_pairs, "pairs_this_rank": n, "epochs": {"rigid": er, "affine": ea, "direct_flow": flow_epochs},
            "ms_total": ms, "voxel_warps_per_s": total_pairs * vox * (er + ea + flow_epochs) / (ms * 1e-3), "scaling": "strong"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-scale", type=float, default=1.0,
                    help="fraction of the README schedule (500 rigid + 200 affine epochs) per e2e step")
    ap.add_argument("--e2e-steps", type=int, default=20)
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
    e0.record()
    prob.run(K, 1e-5, 0.0, 1.0)           # K epochs, ONE C-ABI call: persistent kernel, no per-epoch launch
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    losses = prob.losses[0, W:W + 3].tolist()
    value = world * PAIRS_PER_GPU * vox * K / (ms * 1e-3)
    launches = -(-K // CHUNK_EPOCHS) + 2          # persistent launches + set_contributions_kernel + target_sums_kernel (once per call)

    # ---- end to end through the public API with host buffers ------------------------------
    er, ea = max(1, int(README_EPOCHS[0] * args.e2e_scale)), max(1, int(README_EPOCHS[1] * args.e2e_scale))
    host_m = mov.cpu().pin_memory(); host_t = tgt.cpu().pin_memory()
    reg0 = torch.tensor([0.02, -0.01, 0.03, 0.05, -0.05, 0.02])
    copy_stream = torch.cuda.Stream(device=dev)
    reg0_batch = reg0.repeat(PAIRS_PER_GPU, 1)              # host tensor: no device round trip for the kernel-variant hint
    res_host = torch.empty(PAIRS_PER_GPU, 24, dtype=torch.float32).pin_memory()

    # two resident device buffers per input, refilled alternately from pinned host memory on a side stream: the next
    # batch's H2D copy overlaps the current batch's epochs and no device memory is allocated inside the timed region
    dbuf = [(torch.empty_like(mov), torch.empty_like(tgt)) for _ in range(2)]
    free_ev = [None, None]                  # recorded on the compute stream when a buffer's batch is done
    turn = [0]

    def fetch():
        """H2D of one batch of pairs from pinned host memory into the next device buffer."""
        i = turn[0] % 2
        turn[0] += 1
        m, t = dbuf[i]
        with torch.cuda.stream(copy_stream):
            if free_ev[i] is not None:
                copy_stream.wait_event(free_ev[i])
            m.copy_(host_m, non_blocking=True)
            t.copy_(host_t, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream)
        return m, t, ev, i

    def e2e_step(cur, prefetch):
        """One batch through the public API: Register (batch extension: [N,1,D,H,W] = N independent pairs) rigid ->
        warp -> affine -> thetas to the host."""
        m, t, ev, i = cur
        cs = torch.cuda.current_stream(dev)
        cs.wait_event(ev)
        nxt = fetch() if prefetch else None
        r = tr.Register(mode="rigid", device=dev, weight=[0.0, 1.0, 0.0])
        r.optim(m, t, lr=1e-5, max_epochs=er, reg0=reg0_batch)
        m2 = r(m)
        a = tr.Register(mode="affine", device=dev, weight=[0.0, 1.0, 0.0])
        a.optim(m2, t, lr=1e-5, max_epochs=ea)
        free_ev[i] = torch.cuda.Event(); free_ev[i].record(cs)
        out = torch.cat([r.theta.reshape(PAIRS_PER_GPU, -1), a.theta.reshape(PAIRS_PER_GPU, -1)], 1)
        res_host.copy_(out, non_blocking=True)          # D2H of the step's result into pinned memory
        return res_host, nxt

    _, _ = e2e_step(fetch(), False)                  # warm-up
    torch.cuda.synchronize(dev)
    barrier()
    n_e2e = max(1, args.e2e_steps)
    t_load0 = time.time()
    t0 = time.perf_counter()
    cur = fetch()
    for i in range(n_e2e):
        res, cur = e2e_step(cur, i + 1 < n_e2e)
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    t_load1 = time.time()
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * PAIRS_PER_GPU * vox * (er + ea) * n_e2e / float(dt.item())
    h2d = PAIRS_PER_GPU * 2 * vox * 4
    d2h = PAIRS_PER_GPU * 24 * 4
    clocks = sampler.stop() if rank == 0 else None

    peak, peak_src = _peaks()
    traffic, traffic_src = _traffic()
    us_epoch = ms * 1e3 / K
    kernel = "trb::affine3d_persist_kernel<MSE_ONLY=false, ROT=false> (csrc/affine_persist.cu)"
    roof = roofline_entry(us_epoch, SHAPE, PAIRS_PER_GPU, peak, kernel)
    roof.update({"traffic": traffic * min(K, CHUNK_EPOCHS) if traffic else None, "traffic_per_epoch": traffic, "traffic_source": traffic_src,
                 "peak_source": peak_src, "epochs_per_launch": min(K, CHUNK_EPOCHS),
                 "algorithmic_bytes_per_launch": BYTES_PER_VOXEL_WARP * PAIRS_PER_GPU * vox * min(K, CHUNK_EPOCHS),
                 "kernel_us": us_epoch * min(K, CHUNK_EPOCHS)})
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
        "gpu_launches": launches,
        "gpu_launches_note": "%d cooperative launch(es) of the persistent epoch kernel (<= %d epochs each) + 1 target_sums_kernel "
                             "for the %d timed steps" % (launches - 1, CHUNK_EPOCHS, K),
        "roofline": roof,
        "clocks": clocks,
        "first_losses": losses,
    }

    extra = {}
    if not args.no_extra and world == 1:
        # BASELINE north_star bar: 256^3 on one GPU — first-class roofline of the same kernel
        m256, t256 = make_pair((256, 256, 256), "affine", device=dev)
        us, _ = time_epochs(TF, torch, dev, m256, t256, "affine", ident, 400, (0.0, 1.0), repeats=2)
        line["roofline_256"] = roofline_entry(us, (256, 256, 256), 1, peak, kernel)
        line["roofline_256"]["workload"] = "ONE 256^3 pair (134 MB > L2), affine + NCC, 400 epochs per launch; every epoch ends in a grid-wide hand-over"
        us_r, _ = time_epochs(TF, torch, dev, m256, t256, "rigid", torch.tensor([0.02, -0.01, 0.03, 0.05, -0.05, 0.02], device=dev), 400, (0.0, 1.0))
        extra["single_pair_256^3_rigid"] = {"us_per_epoch": us_r, "iters_per_s": 1e6 / us_r}
        # the reference's own start: torch.manual_seed(0); torch.rand(6) (angles up to 1 rad) — large-rotation path
        torch.manual_seed(0)
        rand0 = torch.rand(6)
        us_l, _ = time_epochs(TF, torch, dev, m256, t256, "rigid", rand0.to(dev), 200, (0.0, 1.0))
        extra["single_pair_256^3_rigid_reference_rand_start"] = {"us_per_epoch": us_l, "ratio_to_small_angle": us_l / us_r,
                                                                  "reg0": rand0.tolist()}
        del m256, t256
        torch.cuda.empty_cache()
        one_m, one_t = mov[:1].contiguous(), tgt[:1].contiguous()
        us, _ = time_epochs(TF, torch, dev, one_m, one_t, "affine", ident, 400, (0.0, 1.0))
        extra["single_pair_192x192x160_L2_resident"] = {"us_per_epoch": us, "iters_per_s": 1e6 / us, "voxel_warps_per_s": vox / (us * 1e-6),
                                                        "algorithmic_GBps": 8.0 * vox / (us * 1e-6) / 1e9}
        # the reference's "criterion given -> MSE only" branch (warpings.py:38-40,125-127) on the headline batch
        us, _ = time_epochs(TF, torch, dev, mov, tgt, "affine", ident, 200, (1.0, 0.0))
        extra["batch_8x192x192x160_mse_only"] = {"us_per_epoch": us, "voxel_warps_per_s": PAIRS_PER_GPU * vox / (us * 1e-6),
                                                 "algorithmic_GBps": 8.0 * PAIRS_PER_GPU * vox / (us * 1e-6) / 1e9,
                                                 "frac": 8.0 * PAIRS_PER_GPU * vox / (us * 1e-6) / 1e9 / peak}
        # theta-Adam (north_star item 3) on the headline batch
        us, _ = time_epochs(TF, torch, dev, mov, tgt, "affine", ident, 200, (0.0, 1.0), optimiser="adam")
        extra["batch_8x192x192x160_ncc_adam"] = {"us_per_epoch": us}
        # the per-epoch-launch kernel of round 1 on the same batch (A/B)
        TF.set_kernel_path("tma")
        us, _ = time_epochs(TF, torch, dev, mov, tgt, "affine", ident, 200, (0.0, 1.0))
        TF.set_kernel_path("auto")
        extra["batch_8x192x192x160_round1_per_epoch_kernel"] = {"us_per_epoch": us}
        # the reference's DEFAULT loss (weights .33/.33/.33 incl. the NMI/KDE term) through the stock Register call: one
        # C-ABI call per optim() (csrc/nmi_src.cu); per-epoch time = difference of a 130- and a 30-epoch call (wall clock)
        def default_loss_us(m_, t_):
            rd = tr.Register(mode="affine", device=dev)
            rd.optim(m_, t_, lr=1e-5, max_epochs=30)        # warm-up with the timed call's own shape (allocator, lazy module loads)
            wall = []
            for ep in (30, 130):
                best = 1e9
                for _ in range(3):      # min of 3: a call's set-up (allocator, value bounds, target moments) varies by milliseconds
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                    rd.optim(m_, t_, lr=1e-5, max_epochs=ep)
                    torch.cuda.synchronize(dev)
                    best = min(best, time.perf_counter() - t0)
                wall.append(best)
            return (wall[1] - wall[0]) / 100 * 1e6
        us = default_loss_us(one_m, one_t)
        extra["single_pair_default_loss_mse+ncc+nmi"] = {"us_per_epoch": us, "voxel_warps_per_s": vox / (us * 1e-6),
                                                         "note": "wall clock through Register, per-epoch slope"}
        us = default_loss_us(mov, tgt)
        extra["batch_8x192x192x160_default_loss_mse+ncc+nmi"] = {"us_per_epoch": us, "voxel_warps_per_s": PAIRS_PER_GPU * vox / (us * 1e-6)}
        # BASELINE configs[2] as the reference runs it: Register(mode='flow') = the attention U-Net (n = 32) + fused head/warp/
        # similarity node at 256^3, MSE + NCC, per-epoch slope of the stock call (wall clock; InstanceNorm / thin convolutions on
        # csrc/instnorm.cu, csrc/thinconv.cu, the rest cuDNN); and the direct per-voxel flow with the smoothness regulariser
        del mov, tgt
        torch.cuda.empty_cache()
        try:
            import torch.nn as nn
            fm, ft = make_pair((256, 256, 256), "flow", device=dev)
            wall = []
            for ep in (2, 2, 6):
                torch.manual_seed(0)
                fr = tr.flow_register((256, 256, 256), mode="bilinear", n=32, lr=1e-3, max_epochs=ep, criterions=[nn.MSELoss(), tr.NCCLoss()],
                                      weights=[0.5, 0.5], stop_crit=-1.0).to(dev)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                fr.optimize(fm, ft, dev, debug=False)
                torch.cuda.synchronize(dev)
                wall.append(time.perf_counter() - t0)
                del fr
            ms = (wall[2] - wall[1]) / 4 * 1e3
            extra["configs2_flow_unet_256^3_mse+ncc"] = {"ms_per_epoch": ms, "voxel_warps_per_s": 256 ** 3 / (ms * 1e-3),
                                                         "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9}
            dp = TF.DirectFlowProblem(fm, ft, 120, optimiser="sgd")
            dp.run(20, 0.05, 0.5, 0.5, 2.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); dp.run(100, 0.05, 0.5, 0.5, 2.0); b.record()
            torch.cuda.synchronize(dev)
            us = a.elapsed_time(b) * 10.0
            extra["configs2_direct_flow_256^3_sgd_mse+ncc+smooth"] = {"us_per_epoch": us, "voxel_warps_per_s": 256 ** 3 / (us * 1e-6),
                                                                      "frac_of_hbm_peak_at_32B_per_voxel": 32.0 * 256 ** 3 / (us * 1e-6) / 1e9 / peak}
            del fm, ft, dp
        except Exception as e:          # keep the headline line even if an extra fails
            extra["configs2_flow"] = {"error": repr(e)}
    if not args.no_extra and world > 1:
        del prob
        torch.cuda.empty_cache()
        try:
            extra["configs3_strong"] = configs3_strong(torch, dist, dev, rank, world)
        except Exception as e:          # keep the headline line even if an extra fails
            extra["configs3_strong"] = {"error": repr(e)}
        del mov, tgt, host_m, host_t
        torch.cuda.empty_cache()
        try:
            extra["sharded_512"] = sharded_512(torch, dist, dev, rank, world)
        except Exception as e:
            extra["sharded_512"] = {"error": repr(e)}
        if world == 2:
            try:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import mgpu_check
                extra["mgpu_check"] = mgpu_check.run_checks(rank, world, dev)
            except Exception as e:
                extra["mgpu_check"] = "FAILED: %r" % (e,)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    if extra:
        line["extra"] = extra
    if not args.no_cpu_baseline and world == 1:
        cb, _ = cpu_reference_leg(3)
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
