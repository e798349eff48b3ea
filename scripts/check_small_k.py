"""GPU check for the opt-in small-K tensor-core path (run with BANET_TC_SMALLK=1): K = 64 / 32 through lm_build_tc6 against the
FP32 SIMT path and the oracle, plus timing of both at 640x480 (the cfg5 sweep of BASELINE.json: tensor cores vs warp-reduce)."""
import os, sys, torch
os.environ.setdefault("BANET_TC_SMALLK", "1")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from banet_b200 import ops, synth
from helpers import O, oracle_level_inputs, rel_fro
for K in (64, 32):
    sc = synth.make_scene(nb=3, H=96, W=128, C=64, K=K, level_ids=(3,), seed=50 + K, device="cpu", dtype=torch.float32)
    lv = sc.levels[0]
    a = oracle_level_inputs(lv)
    Wt = sc.W0 + 0.01 * torch.randn(sc.W0.shape, generator=torch.Generator().manual_seed(1))
    rH, rg, _, rnv = O.normal_equations_structured(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"],
                                                   sc.R0.double(), sc.T0.double(), Wt.double())
    cu = lambda t: t.cuda()
    for grid in (None, lv.grid):
        L = ops.Level(cu(lv.conv1), cu(lv.conv2), cu(lv.intr), cu(lv.p), cu(lv.D), cu(lv.B), grid=grid)
        for prec in (0, 2, 3):
            H, g, rbar, nv = ops.lm_build(L, cu(sc.R0), cu(sc.T0), cu(Wt), precision=prec)
            print(f"K={K} grid={grid is not None} prec={prec}: relH {rel_fro(H, rH):.2e} relg {rel_fro(g, rg.squeeze(-1)):.2e} nvalid ok {torch.equal(nv.cpu().double(), rnv)}")
    big = synth.make_scene(nb=8, H=480, W=640, C=128, K=K, level_ids=(3,), seed=5, device="cuda", dtype=torch.float32)
    bl = big.levels[0]
    L = ops.Level(bl.conv1, bl.conv2, bl.intr, bl.p, bl.D, bl.B, grid=bl.grid)
    for prec in (0, 2):
        for _ in range(2): ops.lm_build(L, big.R0, big.T0, big.W0, precision=prec)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.lm_build(L, big.R0, big.T0, big.W0, precision=prec)
        e1.record(); torch.cuda.synchronize()
        print(f"K={K} 640x480 nb=8 prec={prec}: {e0.elapsed_time(e1) / 5:.3f} ms")
