"""CPU simulation (oracle maths in float64 + emulated tf32 rounding of the two MMA operands) of what each tensor-core precision
scheme does to H, g and the LM step on a bench-like smooth scene: hardware truncation of the basis (old TF32X1), round-to-nearest,
stochastic rounding, exact split (TF32X2).  Run: python scripts/sim_precision_modes.py   (no GPU needed; imports the oracle — a
diagnostic, not part of the product)."""
import sys, os, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import O, oracle_level_inputs, rel_fro
from oracle import ba_oracle as BO
from banet_b200 import synth
torch.manual_seed(0)
def bits(x): return x.float().contiguous().view(torch.int32)
def rna_tf32(x): return ((bits(x) + 0x1000) & ~0x1fff).view(torch.float32).double()
def trunc_tf32(x): return (bits(x) & ~0x1fff).view(torch.float32).double()
def sr_tf32(x, gen):   # stochastic rounding: add uniform 13-bit dither then truncate
    d = torch.randint(0, 0x2000, x.shape, generator=gen, dtype=torch.int32)
    return ((bits(x) + d) & ~0x1fff).view(torch.float32).double()
sc = synth.make_scene(nb=2, H=480, W=640, C=16, K=128, level_ids=(1, 2), seed=1236, device="cpu", dtype=torch.float32)
gen = torch.Generator().manual_seed(5)
lam_base = None
for li, l in enumerate(sc.levels):
    a = oracle_level_inputs(l)
    R, T, W = sc.R0.double(), sc.T0.double(), sc.W0.double()
    conv1, conv2, fx, fy, ox, oy, p, D, B = [a[k] for k in ("conv1","conv2","fx","fy","ox","oy","p","D","B")]
    nb, N, C = conv1.shape
    Dt = D + B @ W
    Rp, x, y, Z, px, py = BO._warp(p, Dt, R, T, fx, fy, ox, oy)
    diff, grad, m = BO._sample_diff_grad(conv1, conv2, px, py, False)
    M = grad.transpose(-1, -2) @ grad; q = grad.transpose(-1, -2) @ diff
    Jc = BO.camera_jacobian_matrix(x, y, Z, fx, fy)
    Hcc = (Jc.transpose(-1, -2) @ M @ Jc).sum(1); gc = (Jc.transpose(-1, -2) @ q).sum(1)
    jd = BO.depth_jacobian_matrix(Rp[:, 0:1], Rp[:, 1:2], Rp[:, 2:3], x, y, Z, fx, fy).unsqueeze(-1)
    v = (Jc.transpose(-1, -2) @ M @ jd).squeeze(-1); s = (jd.transpose(-1, -2) @ M @ jd).reshape(nb, N); t = (jd.transpose(-1, -2) @ q).reshape(nb, N)
    Bf = B.float().double()
    rbar = diff.squeeze(-1).abs().mean(1)
    lam = 1000.0 * 0.01   # fixed stand-in for the lambda MLP
    def assemble(Hdd, Hcd, gd):
        H = torch.cat([torch.cat([Hcc, Hcd], 2), torch.cat([Hcd.transpose(1, 2), Hdd], 2)], 1)
        return H, torch.cat([gc, gd], 1)
    def solve(H, g):
        Hd = H + torch.diag_embed(torch.cat([(torch.diagonal(H, dim1=1, dim2=2)[:, :-1] + 1e-5) * lam, torch.zeros(nb, 1, dtype=torch.float64)], 1))
        return torch.linalg.solve(Hd, g)
    ext = torch.cat([v, t.unsqueeze(-1)], -1)
    def build(A, Rm, E):      # A [nb,N,K] MMA A operand; Rm [nb,N,K]; E [nb,N,7]
        Hdd = A.transpose(1, 2) @ Rm; X = A.transpose(1, 2) @ E
        return assemble(Hdd, X[:, :, :6].transpose(1, 2), X[:, :, 6:7])
    H0, g0 = build(Bf, s.unsqueeze(-1) * Bf, ext); sol0 = solve(H0, g0)
    Arn = rna_tf32(Bf); Asr = sr_tf32(Bf, gen)
    cases = {
      "X2  (A exact, R rna)":            (Bf,  rna_tf32(s.unsqueeze(-1) * Bf),  rna_tf32(ext)),
      "X2s (A exact, R stoch)":          (Bf,  sr_tf32(s.unsqueeze(-1) * Bf, gen),  sr_tf32(ext, gen)),
      "X1  (A trunc, R rna)":            (trunc_tf32(Bf), rna_tf32(s.unsqueeze(-1) * Bf), rna_tf32(ext)),
      "X1R (A rna, R rna(sA))":          (Arn, rna_tf32(s.unsqueeze(-1) * Arn), rna_tf32(ext)),
      "X1S (A stoch, R rna(sA))":        (Asr, rna_tf32(s.unsqueeze(-1) * Asr), rna_tf32(ext)),
      "X1SS(A stoch, R stoch(sA))":      (Asr, sr_tf32(s.unsqueeze(-1) * Asr, gen), sr_tf32(ext, gen)),
      "X1S'(A stoch, R stoch(s*b exact))": (Asr, sr_tf32(s.unsqueeze(-1) * Bf, gen), sr_tf32(ext, gen)),
    }
    print(f"level {l.level} {l.w}x{l.h} N={N}")
    for name, (A, Rm, E) in cases.items():
        H, g = build(A, Rm, E); sol = solve(H, g)
        print(f"  {name:36s} relH={rel_fro(H, H0):.2e} relg={rel_fro(g, g0):.2e}  rel delta={rel_fro(sol, sol0):.2e} (pose {rel_fro(sol[:, :6], sol0[:, :6]):.2e}, depth {rel_fro(sol[:, 6:], sol0[:, 6:]):.2e})")
    H, g = H0.float().double(), g0.float().double(); sol = solve(H, g)
    print(f"  {'fp32-rounded H,g':36s} relH={rel_fro(H, H0):.2e} relg={rel_fro(g, g0):.2e}  rel delta={rel_fro(sol, sol0):.2e}")
