"""GPU: outputs of the bench workload (cfg2: nb pairs, 4 levels x 5 iterations, lambda-MLP) in every precision mode against the
FP32 SIMT path (which the parity tests pin to the oracle)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from banet_b200 import ops, synth, _lib
nb = int(os.environ.get("BANET_NB", "8")); C = K = 128; iters = 5
sc = synth.make_scene(nb=nb, H=bench.H_FULL, W=bench.W_FULL, C=C, K=K, level_ids=bench.LEVEL_IDS, seed=1234 + 2, device="cuda", dtype=torch.float32)
g = torch.Generator().manual_seed(7)
dims = [C, 2 * C, 4 * C, 2 * C, C, 1]
packed = []
for _ in bench.LEVEL_IDS:
    params = [(torch.randn(dims[i], dims[i + 1], generator=g) * (2.0 / dims[i]) ** 0.5, torch.zeros(dims[i + 1])) for i in range(5)]
    packed.append(ops.pack_mlp(params).cuda())
rf = lambda a, b: ((a - b).norm() / b.norm()).item()
def run(prec, fly):
    levels = [ops.Level(l.conv1, l.conv2[..., :C].contiguous() if fly else l.conv2, l.intr, l.p, l.D, l.B, grid=l.grid) for l in sc.levels]
    return ops.lm_run(levels, iters, sc.R0, sc.T0, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0, precision=prec)
R0, T0, W0, st = run(_lib.PREC_FP32_SIMT, False)
print(f"fp32 simt: status {int(st.abs().max())}; moved R by {rf(R0, sc.R0):.2e}, T by {rf(T0, sc.T0):.2e}, |W| {W0.norm().item():.3e}")
modes = [m for m in (("tf32x3", 3), ("tf32x2", 2), ("tf32x1", 1)) if str(m[1]) in os.environ.get("BANET_PRECS", "1,2,3").split(",")]
for name, prec in modes:
    for fly in (False,):
        R, T, W, st = run(prec, fly)
        print(f"{name} fly={int(fly)}: rel-fro vs fp32 simt  R {rf(R, R0):.2e}  T {rf(T, T0):.2e}  W {rf(W, W0):.2e}  status {int(st.abs().max())}")
        if not fly: print("    per pair W:", " ".join(f"{rf(W[i], W0[i]):.1e}" for i in range(nb)))
R, T, W, st = run(_lib.PREC_FP32_SIMT, True)
print(f"fp32 simt fly=1: R {rf(R, R0):.2e}  T {rf(T, T0):.2e}  W {rf(W, W0):.2e}")
