"""Loader of oracle/_ref/libbanet_ref_eqc.so — the reference's own EquationConstruction(+Grad) op kernels, compiled unmodified from
/root/reference/utils.cu (oracle/Makefile).  TEST INFRASTRUCTURE ONLY.  Needs a CUDA device (the op is cuBLAS + two CUDA kernels).

The reference op hoards process-static scratch sized by its first call (utils.cu:210-216), so every distinct shape gets its own
freshly loaded copy of the library."""
import ctypes as C
import os
import shutil
import tempfile

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libbanet_ref_eqc.so")
_instances = {}
_tmpdir = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _instance(shape):
    global _tmpdir
    if shape not in _instances:
        if _tmpdir is None:
            _tmpdir = tempfile.mkdtemp(prefix="banet_ref_")
        path = os.path.join(_tmpdir, "ref_%d_%d_%d_%d.so" % shape)
        shutil.copyfile(LIB_PATH, path)
        lib = C.CDLL(path)
        lib.banet_ref_eqc_fwd.restype = C.c_int
        lib.banet_ref_eqc_fwd.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p] * 2
        lib.banet_ref_eqc_bwd.restype = C.c_int
        lib.banet_ref_eqc_bwd.argtypes = [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p] * 3
        _instances[shape] = lib
    return _instances[shape]


def _c(t):
    assert t.is_cuda and t.dtype == torch.float32
    return t.contiguous()


def equation_construction(J, G, d):
    """reference op `EquationConstruction` (utils.cu:219-417): J [nb,N,2,P], G [nb,N,C,2], d [nb,N,C,1] (cuda fp32) -> AtA, Atb."""
    J, G, d = _c(J), _c(G), _c(d)
    nb, N, _, P = J.shape; Cc = G.shape[2]
    AtA = torch.empty(nb, P, P, device=J.device); Atb = torch.empty(nb, P, 1, device=J.device)
    torch.cuda.synchronize()
    rc = _instance((nb, N, Cc, P)).banet_ref_eqc_fwd(J.data_ptr(), G.data_ptr(), d.data_ptr(), nb, N, Cc, P, AtA.data_ptr(), Atb.data_ptr())
    if rc:
        raise RuntimeError(f"banet_ref_eqc_fwd failed ({rc})")
    return AtA, Atb


def equation_construction_grad(J, G, d, left_grad, right_grad):
    """reference op `EquationConstructionGrad` (utils.cu:465-694); the forward op must have run for this shape (it owns the scratch)."""
    J, G, d, lg, rg = _c(J), _c(G), _c(d), _c(left_grad), _c(right_grad)
    nb, N, _, P = J.shape; Cc = G.shape[2]
    dJ, dG, dd = torch.empty_like(J), torch.empty_like(G), torch.empty_like(d)
    torch.cuda.synchronize()
    rc = _instance((nb, N, Cc, P)).banet_ref_eqc_bwd(J.data_ptr(), G.data_ptr(), d.data_ptr(), lg.data_ptr(), rg.data_ptr(), nb, N, Cc, P,
                                                     dJ.data_ptr(), dG.data_ptr(), dd.data_ptr())
    if rc:
        raise RuntimeError(f"banet_ref_eqc_bwd failed ({rc})")
    return dJ, dG, dd
