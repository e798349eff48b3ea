#!/usr/bin/env python
"""bench.py — LM iterations/sec of the BA layer's inner loop (BASELINE.json metric) on N B200s.

A "step" is one whole coarse-to-fine solve of BASELINE config 2 on every rank's shard:
nb=32 frame-pairs per GPU, dense levels 80x60 -> 640x480 (ΣN = 408 000 points/pair), C=128 feature
channels, K=128 depth bases, 5 LM iterations per level (20 iterations), lambda-MLP in the loop.
Frame-pairs are independent, so ranks hold disjoint shards (weak scaling) and the only collective is
one all-gather of the solved (R,T,W) at the end of each step.

value      pair-iterations/s  = (total pairs) * 20 / t_step      inputs resident in HBM
e2e        the same through the public API with HOST buffers: pinned host -> device copies of every
           level tensor + the solve + device -> host read of (R,T,W), all inside the timed region
roofline   dominant kernel lm_build_kernel: algorithmic bytes 4*N*(2C+K+4)+4*(P^2+P+C) per pair
            / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
cpu_baseline / --impl reference
           the reference cannot execute here (TF-1.x / python2 / TF headers absent), so the reference arm is
           the oracle's reference-faithful materialised restatement (J,G,d tensors + batched matmul chain +
           LU solve, torch-CPU fp32, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

METRIC = "LM iters/sec (frame-pair LM iterations, 640x480x4-scale, K=128)"
UNIT = "pair-iters/s"
LEVEL_IDS = (0, 1, 2, 3)
H_FULL, W_FULL = 480, 640


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nb", type=int, default=32, help="frame-pairs per GPU")
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--bases", type=int, default=128)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-chunks", type=int, default=4, help="pair chunks of the host pipeline (copy of chunk k+1 overlaps the solve of chunk k)")
    ap.add_argument("--layout", default="concat", choices=["concat", "f2"],
                    help="conv2 in HBM: 'concat' = [F2|gx|gy] (3C, the reference's BundleIteration boundary), 'f2' = F2 only, gradients on the fly")
    ap.add_argument("--no-precision-check", action="store_true")
    ap.add_argument("--e2e-boundary", default="features", choices=["features", "concat"],
                    help="host buffers of the e2e leg: 'features' = the layer boundary of the reference's BundleResize (bundlenet.py:376-399): "
                         "feature maps in, [F2|gx|gy] derived on the device each step (banet_grad_fixed_concat); 'concat' = the 3C tensor itself")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "tf32x1", "tf32x2", "tf32x3"],
                    help="contraction path of the build kernel: auto = tensor cores (tcgen05 tf32 split-A) when K=128, else fp32 SIMT")
    return ap.parse_args()


def algorithmic_bytes_per_pair_iter(N, C, K):
    """dense level, conv1 + each F2 texel once + B + ray/depth + outputs."""
    P = 6 + K
    return 4 * N * (2 * C + K + 4) + 4 * (P * P + P + C)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.index = index; self.lines = []; self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_reference_sample(C, K, reps, seed=1234):
    """Times the oracle's materialised BundleIteration (reference-faithful: J,G,d tensors, matmul chain, LU) on ONE
    pair at the 160x120 level (N = 19 200) in fp32 with all host threads; returns seconds per pair-iteration there."""
    from oracle import ba_oracle as O
    from banet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc = synth.make_scene(nb=1, H=H_FULL, W=W_FULL, C=C, K=K, level_ids=(1,), seed=seed, dtype=torch.float32, device="cpu")
    lv = sc.levels[0]
    fx, fy, ox, oy = lv.intr_tiled()
    mlp = O.init_lambda_mlp(C, dtype=torch.float32)
    times = []
    with torch.no_grad():
        for i in range(reps + 1):
            t0 = time.perf_counter()
            O.bundle_iteration(lv.conv1, lv.conv2, fx, fy, ox, oy, lv.p, lv.D, lv.B, sc.R0, sc.T0, sc.W0, mlp)
            times.append(time.perf_counter() - t0)
    return times[1:], lv.N, cores        # first call is a warm-up


def pixels_per_pair_iter():
    """A pair-iteration of the 4-level workload touches ΣN/4 points on average."""
    tot = sum((H_FULL // 2 ** (3 - l)) * (W_FULL // 2 ** (3 - l)) for l in LEVEL_IDS)
    return tot / len(LEVEL_IDS)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    reps = max(1, args.steps)
    times, n_sample, cores = cpu_reference_sample(args.channels, args.bases, reps)
    t = sorted(times)[len(times) // 2]
    # pair-iterations/s at the workload's mean level size, extrapolated linearly in points (stated)
    value = (n_sample / pixels_per_pair_iter()) / t
    sample = (f"1 pair x 1 BundleIteration at 160x120 (N={n_sample}), C={args.channels}, K={args.bases}, fp32, "
              f"median of {len(times)} runs, extrapolated linearly in points to the mean level size")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": reps,
            "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "reference_arm": "oracle port (TF-1.x reference cannot execute here)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(args):
    return (f"cfg2: nb={args.nb}/GPU frame-pairs, dense levels 80x60..640x480, C={args.channels}, K={args.bases}, "
            f"{args.iters} LM iters/level, lambda-MLP")


# ----------------------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    from banet_b200 import ops, synth, _lib
    from banet_b200 import dist as bdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()
    if world > 1:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=dev)

    C, K, nb, iters = args.channels, args.bases, args.nb, args.iters
    sc = synth.make_scene(nb=nb, H=H_FULL, W=W_FULL, C=C, K=K, level_ids=LEVEL_IDS, seed=1234 + 2 + 1000 * rank,
                          device=dev, dtype=torch.float32)
    if args.layout == "f2":            # keep only the feature third of conv2 (the gradients are recomputed on the fly by the kernel)
        for l in sc.levels:
            l.conv2 = l.conv2[..., :C].contiguous()
    levels = [ops.Level(l.conv1, l.conv2, l.intr, l.p, l.D, l.B, grid=l.grid) for l in sc.levels]
    PREC = {"auto": _lib.PREC_AUTO, "fp32": _lib.PREC_FP32_SIMT, "tf32x1": _lib.PREC_TF32X1, "tf32x2": _lib.PREC_TF32X2,
            "tf32x3": _lib.PREC_TF32X3}[args.precision]
    g = torch.Generator().manual_seed(7)
    dims = [C, 2 * C, 4 * C, 2 * C, C, 1]
    packed = []
    for _ in LEVEL_IDS:       # he-normal lambda-MLP, seed 7 (reference bundlenet.py:105)
        params = [(torch.randn(dims[i], dims[i + 1], generator=g) * (2.0 / dims[i]) ** 0.5, torch.zeros(dims[i + 1])) for i in range(5)]
        packed.append(ops.pack_mlp(params).to(dev))
    ws = torch.empty(ops.lm_run_workspace_bytes(levels, PREC), dtype=torch.uint8, device=dev)
    n_levels, total_iters = len(levels), len(levels) * iters

    def step():
        R, T, W, status = ops.lm_run(levels, iters, sc.R0, sc.T0, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0, workspace=ws,
                                     precision=PREC)
        if world > 1:
            return bdist.all_gather_solution(R, T, W), status
        return (R, T, W), status

    def barrier():
        if world > 1:
            td.barrier(device_ids=[local])
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        out, status = step()
    barrier()
    assert int(status.abs().max()) == 0, "solver reported a non-SPD / non-finite system"

    # ---- accuracy of the timed precision mode on THIS workload: outputs against the FP32 SIMT path (pinned to the oracle by tests/) ----
    precision_check = None
    if rank == 0 and not args.no_precision_check and PREC != _lib.PREC_FP32_SIMT:
        rf = lambda a, b: float(((a - b).norm() / b.norm()).item())
        R1, T1, W1, _ = ops.lm_run(levels, iters, sc.R0, sc.T0, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0, precision=PREC)
        R0_, T0_, W0_, _ = ops.lm_run(levels, iters, sc.R0, sc.T0, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0,
                                     precision=_lib.PREC_FP32_SIMT)
        fin = sc.levels[-1]                      # the layer's depth output D + B.W (reference bundlenet.py:397) at the finest level
        d1 = ops.depth_compose(fin.D.reshape(nb, -1), fin.B, W1); d0 = ops.depth_compose(fin.D.reshape(nb, -1), fin.B, W0_)
        errs = {"R": rf(R1, R0_), "T": rf(T1, T0_), "depth": rf(d1, d0), "W": rf(W1, W0_)}
        precision_check = {"vs": "fp32_simt path, same inputs, all levels x iterations", "rel_fro": errs, "tolerance": 1e-4,
                           "ok": max(errs["R"], errs["T"], errs["depth"]) < 1e-4,
                           "note": "north-star tolerance is on the pose / depth outputs; W (depth-basis coefficients) is reported as well"}
        del d1, d0
        del R1, T1, W1, R0_, T0_, W0_
    barrier()

    sampler = ClockSampler(local); sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out, status = step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    if world > 1:
        tms = torch.tensor([ms], device=dev, dtype=torch.float64)
        td.all_reduce(tms, op=td.ReduceOp.MAX)
        ms = float(tms.item())
    ms_per_step = ms / args.steps
    value = world * nb * total_iters / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (lm_build_kernel), per level, CUDA events on the launch stream ----
    peak, peak_kind = measured_peaks()
    per_level = []
    for lv, sl in zip(levels, sc.levels):
        for _ in range(2):
            ops.lm_build(lv, sc.R0, sc.T0, sc.W0, precision=PREC)
        reps = 5
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            ops.lm_build(lv, sc.R0, sc.T0, sc.W0, precision=PREC)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / reps * 1e-3
        by = nb * algorithmic_bytes_per_pair_iter(sl.N, C, K)
        by3c = by + nb * 4 * sl.N * 2 * C          # conv2 read as the reference lays it out: [F2|gx|gy] = 3C channels per texel
        per_level.append({"level": f"{sl.w}x{sl.h}", "ms": t * 1e3, "alg_bytes": by, "gbs": by / t / 1e9, "gbs_3c_layout": by3c / t / 1e9})
    top = per_level[-1]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "lm_build_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    kname = "lm_build_kernel" if PREC == _lib.PREC_FP32_SIMT else "lm_build_tc6_kernel"
    roofline = {"bound": "hbm", "kernel": f"{kname} (+lm_reduce_kernel) @640x480", "achieved": top["gbs"], "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": top["gbs"] / peak, "frac_3c_layout": top["gbs_3c_layout"] / peak,
                "traffic": traffic, "per_level": per_level,
                "all_levels_gbs": sum(p["alg_bytes"] for p in per_level) / sum(p["ms"] * 1e-3 for p in per_level) / 1e9}

    # ---- e2e: host buffers -> device -> solve -> host, through the public API -------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(args, sc, levels, packed, ws, world, local, dev, total_iters, PREC)
        except Exception as ex:      # e.g. not enough pinnable host memory: report, do not fake
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)[:200]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times, n_sample, cores = cpu_reference_sample(C, K, 4)
        t = sorted(times)[len(times) // 2]
        cpu_baseline = {"value": (n_sample / pixels_per_pair_iter()) / t, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"oracle materialised BundleIteration, 1 pair at 160x120 (N={n_sample}), fp32, median of {len(times)}, "
                                  f"{t:.2f} s each, extrapolated linearly in points"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": workload_name(args), "global_pairs": world * nb, "lm_iterations_per_step": total_iters,
                           "batch_iters_per_s": total_iters / (ms_per_step * 1e-3), "precision": args.precision,
                           "conv2_layout": "[F2|gx|gy] (3C channels, the reference's BundleIteration boundary)" if args.layout == "concat"
                                           else "F2 only (C channels); the kernel recomputes gx, gy on the fly (reference grad_fixed, bundlenet.py:92-100)",
                           "l2": "inputs (~33 GB/GPU) far exceed the 126 MB L2; no flush needed",
                           "parallelism": f"pairs sharded over {world} GPU(s), one all-gather of (R,T,W) per step"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps * (1 + total_iters * 5), "precision_check": precision_check,
                "roofline": roofline, "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


def run_e2e(args, sc, levels, packed, ws, world, local, dev, total_iters, prec):
    """The call a user with host-resident inputs makes: banet_b200.host_pipeline.HostSolver (pair chunks, H2D of chunk k+1
    overlapping the solve of chunk k); every step copies every level tensor host->device and the result device->host."""
    from banet_b200 import dist as bdist
    from banet_b200.host_pipeline import HostSolver
    C = args.channels
    feat = args.e2e_boundary == "features" and args.layout == "concat"
    host = []
    for l in sc.levels:
        tens = {"grid": l.grid}
        for name in ("conv1", "conv2", "intr", "p", "D", "B"):
            t = getattr(l, name)
            if name == "conv2" and feat:
                t = t[..., :C].contiguous()
            ht = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
            ht.copy_(t)
            tens[name] = ht
        host.append(tens)
    solver = HostSolver(host, derive_gradients=feat, chunks=args.e2e_chunks, device=dev, precision=prec)
    hR = sc.R0.cpu().pin_memory(); hT = sc.T0.cpu().pin_memory(); hW = sc.W0.cpu().pin_memory()
    h2d = solver.h2d_bytes + (hR.numel() + hT.numel() + hW.numel()) * 4
    oR = torch.empty_like(hR).pin_memory(); oT = torch.empty_like(hT).pin_memory(); oW = torch.empty_like(hW).pin_memory()
    d2h = (oR.numel() + oT.numel() + oW.numel()) * 4 * world

    def e2e_step():
        R, T, W, status = solver.solve(hR, hT, hW, args.iters, mlp_packed=packed, l2_regularizer_base=1000.0,
                                       out=None if world > 1 else (oR, oT, oW))
        if world > 1:
            R, T, W = bdist.all_gather_solution(R, T, W)
            return R.cpu(), T.cpu(), W.cpu()
        torch.cuda.current_stream().synchronize()
        return oR, oT, oW

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as td
        td.barrier(device_ids=[local])
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        e2e_step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms], device=dev, dtype=torch.float64)
        td.all_reduce(tms, op=td.ReduceOp.MAX); ms = float(tms.item())
    per = ms / args.e2e_steps * 1e-3
    return {"value": world * args.nb * total_iters / per, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
            "d2h_bytes_per_step": d2h, "ms_per_step": per * 1e3, "steps": args.e2e_steps,
            "boundary": ("feature maps (C channels) + conv1, p, D, B, intr in pinned host memory; [F2|gx|gy] derived on the device every step "
                         "(banet_grad_fixed_concat), as the reference's BundleResize does (bundlenet.py:386-389)") if feat else
                        "every level tensor, conv2 as the 3C [F2|gx|gy] tensor, in pinned host memory",
            "api": f"banet_b200.host_pipeline.HostSolver(chunks={args.e2e_chunks}).solve",
            "note": "pinned host -> device copy of every level tensor + solve + device -> host of (R,T,W) per step; copies of pair-chunk k+1 "
                    "overlap the solve of chunk k"}


if __name__ == "__main__":
    main()
