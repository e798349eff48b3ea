"""CPU simulation of WHOLE solves (oracle maths in float64, tf32 rounding of the two MMA operands emulated) to compare one-pass
rounding schemes against the exact-split TF32X2 scheme over many LM iterations: round-to-nearest basis (same perturbation every
iteration), stochastic rounding re-drawn per iteration, and antithetic dither (iteration 2k+1 uses the complement of iteration 2k).
Run: python scripts/sim_precision_solve.py   (no GPU; a diagnostic that imports the oracle, not part of the product)."""
import os, sys, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import oracle_level_inputs, rel_fro
from oracle import ba_oracle as BO
from banet_b200 import synth

def bits(x): return x.float().contiguous().view(torch.int32)
def rna(x): return ((bits(x) + 0x1000) & ~0x1fff).view(torch.float32).double()
def trunc(x): return (bits(x) & ~0x1fff).view(torch.float32).double()
def dith(x, d): return ((bits(x) + d) & ~0x1fff).view(torch.float32).double()

NB, C, K = 4, 16, 128
LAM = float(os.environ.get("LAM", "1e3")); ITERS = int(os.environ.get("ITERS", "5"))
sc = synth.make_scene(nb=NB, H=480, W=640, C=C, K=K, level_ids=(0, 1), seed=1236, device="cpu", dtype=torch.float32)

def iteration(a, R, T, W, scheme, it, gen):
    conv1, conv2, fx, fy, ox, oy, p, D, B = [a[k] for k in ("conv1", "conv2", "fx", "fy", "ox", "oy", "p", "D", "B")]
    nb, N, _ = conv1.shape
    Rp, x, y, Z, px, py = BO._warp(p, D + B @ W, R, T, fx, fy, ox, oy)
    diff, grad, m = BO._sample_diff_grad(conv1, conv2, px, py, False)
    M = grad.transpose(-1, -2) @ grad; q = grad.transpose(-1, -2) @ diff
    Jc = BO.camera_jacobian_matrix(x, y, Z, fx, fy)
    Hcc = (Jc.transpose(-1, -2) @ M @ Jc).sum(1); gc = (Jc.transpose(-1, -2) @ q).sum(1)
    jd = BO.depth_jacobian_matrix(Rp[:, 0:1], Rp[:, 1:2], Rp[:, 2:3], x, y, Z, fx, fy).unsqueeze(-1)
    v = (Jc.transpose(-1, -2) @ M @ jd).squeeze(-1); s = (jd.transpose(-1, -2) @ M @ jd).reshape(nb, N); t = (jd.transpose(-1, -2) @ q).reshape(nb, N)
    ext = torch.cat([v, t.unsqueeze(-1)], -1)
    Bf = B.float().double()
    if scheme == "exact": A, Rm, E = Bf, s.unsqueeze(-1) * Bf, ext
    elif scheme == "x2": A, Rm, E = Bf, rna(s.unsqueeze(-1) * Bf), rna(ext)
    elif scheme == "x1trunc": A = trunc(Bf); Rm, E = rna(s.unsqueeze(-1) * Bf), rna(ext)
    else:
        if scheme == "x1rna": A = rna(Bf)
        elif scheme == "x1stoch": A = dith(Bf, torch.randint(0, 0x2000, Bf.shape, generator=gen, dtype=torch.int32))
        elif scheme == "x1anti":
            g2 = torch.Generator().manual_seed(1000 + it // 2)
            d = torch.randint(0, 0x2000, Bf.shape, generator=g2, dtype=torch.int32)
            A = dith(Bf, d if it % 2 == 0 else 0x1fff - d)
        Rm, E = rna(s.unsqueeze(-1) * A), rna(ext)
    Hdd = A.transpose(1, 2) @ Rm; X = A.transpose(1, 2) @ E
    Hcd = X[:, :, :6].transpose(1, 2); gd = X[:, :, 6:7]
    H = torch.cat([torch.cat([Hcc, Hcd], 2), torch.cat([Hcd.transpose(1, 2), Hdd], 2)], 1); g = torch.cat([gc, gd], 1)
    H = torch.tril(H) + torch.tril(H, -1).transpose(1, 2)            # the reduce kernel mirrors the lower triangle
    Hd = H + torch.diag_embed(torch.cat([(torch.diagonal(H, dim1=1, dim2=2)[:, :-1] + 1e-5) * LAM, torch.zeros(nb, 1, dtype=torch.float64)], 1))
    sol = torch.linalg.solve(Hd, g)
    Rn, Tn = BO._update(sol[:, :6, :], R, T, BO.IterOptions())
    return Rn, Tn, W + sol[:, 6:, :]

def solve(scheme):
    R, T, W = sc.R0.double(), sc.T0.double(), sc.W0.double()
    gen = torch.Generator().manual_seed(99); it = 0
    for l in sc.levels:
        a = oracle_level_inputs(l)
        for _ in range(ITERS):
            R, T, W = iteration(a, R, T, W, scheme, it, gen); it += 1
    return R, T, W

ref = solve("exact")
print(f"levels {[ (l.w, l.h) for l in sc.levels]}, {ITERS} iterations each, lambda {LAM:g}; |W| {ref[2].norm():.3f}")
for scheme in ("x2", "x1trunc", "x1rna", "x1stoch", "x1anti"):
    R, T, W = solve(scheme)
    per = " ".join(f"{rel_fro(W[i], ref[2][i]):.1e}" for i in range(NB))
    print(f"{scheme:8s} rel-fro vs exact: R {rel_fro(R, ref[0]):.1e} T {rel_fro(T, ref[1]):.1e} W {rel_fro(W, ref[2]):.1e}   per pair W: {per}")
