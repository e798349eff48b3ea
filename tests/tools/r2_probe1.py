"""Round-2 GPU probe 1: (a) whole-solve accuracy of every precision policy against the float64 oracle on a cfg2-shaped case
(4 levels, C=K=128, lambda-MLP, 5 iterations per level, nb=2; both conv2 layouts); (b) generation 6 vs 7 timing at 640x480, nb=32."""
import json, os, sys, time, statistics, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from banet_b200 import ops, synth, _lib
from helpers import O, oracle_level_inputs, rel_fro
out = {}
what = os.environ.get("PROBE", "acc,time")

if "acc" in what:
    nb = 2
    sc = synth.make_scene(nb=nb, H=480, W=640, C=128, K=128, level_ids=(0, 1, 2, 3), seed=1236, device="cuda", dtype=torch.float32)
    mlps = [O.init_lambda_mlp(128, seed=7 + l.level, dtype=torch.float32) for l in sc.levels]
    packed = [ops.pack_mlp(m).cuda() for m in mlps]
    t0 = time.time()
    olv = []
    for l, m in zip(sc.levels, mlps):
        class _L: pass
        cl = _L(); cl.conv1, cl.conv2, cl.intr, cl.p, cl.D, cl.B = [t.cpu() for t in (l.conv1, l.conv2, l.intr, l.p, l.D, l.B)]
        cl.N = l.N; cl.intr_tiled = lambda cl=cl: tuple(cl.intr[:, i:i + 1].expand(-1, cl.N).contiguous() for i in range(4))
        a = oracle_level_inputs(cl)
        olv.append(O.LevelInputs(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"], [(w.double(), b.double()) for w, b in m]))
    oR, oT, oW = O.lm_solve_structured(olv, 5, sc.R0.cpu().double(), sc.T0.cpu().double(), sc.W0.cpu().double())
    fin = olv[-1]; oDepth = fin.D + fin.B @ oW
    print(f"oracle fp64 solve: {time.time() - t0:.1f} s", flush=True)
    acc = {}
    for layout in ("3c", "f2"):
        levels = [ops.Level(l.conv1, l.conv2 if layout == "3c" else l.conv2[..., :128].contiguous(), l.intr, l.p, l.D, l.B, grid=l.grid) for l in sc.levels]
        for name, prec in (("fp32", 0), ("x3", 3), ("x2", 2), ("x1", 1), ("levelwise", 4), ("auto", -1)):
            R, T, W, st = ops.lm_run(levels, 5, sc.R0, sc.T0, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0, precision=prec)
            d = sc.levels[-1].D.cpu().double() + sc.levels[-1].B.cpu().double() @ W.cpu().double()
            e = {"R": rel_fro(R, oR), "T": rel_fro(T, oT), "W": rel_fro(W, oW), "depth": rel_fro(d, oDepth), "status": int(st.abs().max()),
                 "W_worst_pair": max(rel_fro(W[b], oW[b]) for b in range(nb))}
            acc[f"{layout}/{name}"] = e
            print(f"{layout}/{name}: " + " ".join(f"{k}={v:.2e}" if isinstance(v, float) else f"{k}={v}" for k, v in e.items()), flush=True)
    out["accuracy_vs_fp64_oracle"] = acc
    del sc, levels, olv
    torch.cuda.empty_cache()

if "time" in what:
    nb = int(os.environ.get("BANET_NB", "32"))
    sc = synth.make_scene(nb=nb, H=480, W=640, C=128, K=128, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
    lv = sc.levels[0]
    f2 = lv.conv2[..., :128].contiguous()
    L3 = ops.Level(lv.conv1, lv.conv2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
    Lf = ops.Level(lv.conv1, f2, lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
    cases = [("gen6 3c x1", L3, 1, dict(tc_generation=6)), ("gen6 3c x2", L3, 2, dict(tc_generation=6)), ("gen6 f2 x1", Lf, 1, dict(tc_generation=6)),
             ("gen6 f2 x2", Lf, 2, dict(tc_generation=6))]
    for band in (1, 2, 4, 8):
        cases.append((f"gen7 f2 x1 band{band}", Lf, 1, dict(tc_generation=7, tc7_band_rows=band)))
    cases.append(("gen7 f2 x2 band4", Lf, 2, dict(tc_generation=7, tc7_band_rows=4)))
    cases.append(("gen7 f2 x1 direct", Lf, 1, dict(tc_generation=7, tc7_force_direct=True)))
    res = {c[0]: [] for c in cases}
    ref = {}
    for name, L, prec, tun in cases:
        _lib.set_tuning(**tun)
        for _ in range(2): r = ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
        ref[name] = r
    torch.cuda.synchronize()
    for rnd in range(4):
        for name, L, prec, tun in cases:
            _lib.set_tuning(**tun)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
            e0.record()
            for _ in range(4): ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec)
            e1.record(); torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / 4)
    _lib.set_tuning()
    Hs = ops.lm_build(Lf, sc.R0, sc.T0, sc.W0, precision=0)[0]
    tim = {}
    for name, *_ in cases:
        v = res[name]
        tim[name] = {"min_ms": min(v), "med_ms": statistics.median(v), "relH_vs_simt": rel_fro(ref[name][0], Hs)}
        print(f"{name:24s} min {min(v):7.3f} med {statistics.median(v):7.3f} ms  relH {tim[name]['relH_vs_simt']:.2e}", flush=True)
    out["time_640x480_nb%d" % nb] = tim

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = os.environ.get("PROBE_TAG", "r2_probe1")
json.dump(out, open(os.path.join(ROOT, "gpurun_out", tag + ".json"), "w"), indent=1)
