"""GPU diagnostic: per level of the bench workload, H / g / LM step of each tensor-core mode against the FP32 SIMT build,
at the start point and at the converged point of the FP32 run."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from banet_b200 import ops, synth, _lib
nb = int(os.environ.get("BANET_NB", "4")); C = K = 128
sc = synth.make_scene(nb=nb, H=bench.H_FULL, W=bench.W_FULL, C=C, K=K, level_ids=bench.LEVEL_IDS, seed=1234 + 2, device="cuda", dtype=torch.float32)
g = torch.Generator().manual_seed(7)
dims = [C, 2 * C, 4 * C, 2 * C, C, 1]
packed = []
for _ in bench.LEVEL_IDS:
    params = [(torch.randn(dims[i], dims[i + 1], generator=g) * (2.0 / dims[i]) ** 0.5, torch.zeros(dims[i + 1])) for i in range(5)]
    packed.append(ops.pack_mlp(params).cuda())
levels = [ops.Level(l.conv1, l.conv2, l.intr, l.p, l.D, l.B, grid=l.grid) for l in sc.levels]
Rf, Tf, Wf, _ = ops.lm_run(levels, 5, sc.R0, sc.T0, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0, precision=0)
rf = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
def step(H, gg, lam):
    H = H.double().cpu(); gg = gg.double().cpu().unsqueeze(-1); lam = lam.double().cpu()
    d = torch.diagonal(H, dim1=1, dim2=2)
    Hd = H + torch.diag_embed(torch.cat([(d[:, :-1] + 1e-5) * lam[:, None], torch.zeros(H.shape[0], 1, dtype=torch.float64)], 1))
    return torch.linalg.solve(Hd, gg)
for name, (R, T, W) in (("start", (sc.R0, sc.T0, sc.W0)), ("fp32 end", (Rf, Tf, Wf))):
    for lv, sl, pk in zip(levels, sc.levels, packed):
        H0, g0, rbar, nv = ops.lm_build(lv, R, T, W, precision=0)
        lam = ops.lm_lambda(rbar, sl.N, pk, 1000.0)
        d0 = step(H0, g0, lam)
        out = f"{name:8s} {sl.w}x{sl.h} lam~{lam.mean().item():.2e} |g|/|g|start.. |delta|={d0.norm().item():.2e}:"
        for prec in (3, 2, 1):
            H, gg, _, _ = ops.lm_build(lv, R, T, W, precision=prec)
            d = step(H, gg, lam)
            out += f"  X{prec}: relH {rf(H, H0):.1e} relg {rf(gg, g0):.1e} reld {rf(d, d0):.1e}"
        print(out)
