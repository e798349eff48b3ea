"""Writes ref_eqc.npz: inputs and outputs of the reference's OWN compiled op kernels (EquationConstruction + Grad, utils.cu, built
unmodified by oracle/Makefile) on a B200.  Run on the GPU box:  python tests/golden/gen_ref_eqc_golden.py [outdir]
(outdir defaults to gpurun_out/, which the gpurun client copies back; the file is then committed under tests/golden/)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_lib      # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
g = torch.Generator().manual_seed(4242)
nb, N, C, P = 2, 96, 16, 38
J = torch.randn(nb, N, 2, P, generator=g); G = torch.randn(nb, N, C, 2, generator=g); d = torch.randn(nb, N, C, 1, generator=g)
lg = torch.randn(nb, P, P, generator=g); rg = torch.randn(nb, P, 1, generator=g)
A, b = ref_lib.equation_construction(J.cuda(), G.cuda(), d.cuda())
dJ, dG, dd = ref_lib.equation_construction_grad(J.cuda(), G.cuda(), d.cuda(), lg.cuda(), rg.cuda())
os.makedirs(out_dir, exist_ok=True)
np.savez_compressed(os.path.join(out_dir, "ref_eqc.npz"), in_J=J.numpy(), in_G=G.numpy(), in_d=d.numpy(), in_left_grad=lg.numpy(), in_right_grad=rg.numpy(),
                    out_AtA=A.cpu().numpy(), out_Atb=b.cpu().numpy(), out_dJ=dJ.cpu().numpy(), out_dG=dG.cpu().numpy(), out_dd=dd.cpu().numpy(),
                    meta=np.array([torch.cuda.get_device_name(0), torch.version.cuda]))
print("wrote", os.path.join(out_dir, "ref_eqc.npz"))
