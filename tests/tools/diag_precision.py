"""Diagnostic (GPU): where does the output error of one LM iteration come from?  Run under gpurun."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from helpers import O, scene_case, oracle_level_inputs, to_cuda32, rel_fro
from banet_b200 import ops

sc = scene_case(nb=2, H=96, W=128, C=64, K=128, level_ids=(2, 3), seed=91, dtype=torch.float32)
lam = 0.05
for li, l in enumerate(sc.levels):
    a = oracle_level_inputs(l)
    R, T, W = sc.R0.double(), sc.T0.double(), sc.W0.double() + (0.01 if li else 0.0)
    opts = O.IterOptions(lambda_override=torch.full((2,), lam, dtype=torch.float64))
    Rn, Tn, Wn, aux = O.bundle_iteration(a["conv1"], a["conv2"], a["fx"], a["fy"], a["ox"], a["oy"], a["p"], a["D"], a["B"], R, T, W, None, opts, return_aux=True)
    H64, g64, sol64 = aux["AtA"], aux["Atb"], aux["solution"]
    Hd = H64 + torch.diag_embed(torch.cat([(torch.diagonal(H64, dim1=1, dim2=2)[:, :-1] + 1e-5) * lam, torch.zeros(2, 1, dtype=torch.float64)], 1))
    print(f"level {l.level}: N={l.N} cond(H damped)={torch.linalg.cond(Hd).tolist()}")
    lvl = ops.Level(to_cuda32(l.conv1), to_cuda32(l.conv2), to_cuda32(l.intr), to_cuda32(l.p), to_cuda32(l.D), to_cuda32(l.B))
    for prec in (0, 2, 1):
        H, g, rbar, nv = ops.lm_build(lvl, to_cuda32(R), to_cuda32(T), to_cuda32(W), precision=prec)
        lamt = torch.full((2,), lam, device="cuda")
        Rg, Tg, Wg, delta, st = ops.lm_solve_update(H, g, lamt, to_cuda32(R), to_cuda32(T), to_cuda32(W))
        # solve the GPU's H,g in float64 on the host: separates build error from solve error
        Hh, gh = H.double().cpu(), g.double().cpu().unsqueeze(-1)
        Hhd = Hh + torch.diag_embed(torch.cat([(torch.diagonal(Hh, dim1=1, dim2=2)[:, :-1] + 1e-5) * lam, torch.zeros(2, 1, dtype=torch.float64)], 1))
        sol_h = torch.linalg.solve(Hhd, gh)
        # oracle H,g rounded to fp32 then solved in fp64: the error floor of ANY fp32-output build
        H32, g32 = H64.float().double(), g64.float().double()
        H32d = H32 + torch.diag_embed(torch.cat([(torch.diagonal(H32, dim1=1, dim2=2)[:, :-1] + 1e-5) * lam, torch.zeros(2, 1, dtype=torch.float64)], 1))
        sol_r = torch.linalg.solve(H32d, g32)
        print(f"  prec={prec}: relH={rel_fro(H, H64):.2e} relg={rel_fro(g, g64.squeeze(-1)):.2e} | delta: gpu={rel_fro(delta, sol64.squeeze(-1)):.2e} "
              f"host-solve-of-gpu-H={rel_fro(sol_h, sol64):.2e} fp32-rounded-oracle-H={rel_fro(sol_r, sol64):.2e} | W'={rel_fro(Wg, Wn):.2e} T'={rel_fro(Tg, Tn):.2e}")
