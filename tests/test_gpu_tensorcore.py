"""GPU tests of the tcgen05 / TMA building blocks (SWIZZLE_128B MN-major operands, kind::tf32) against fp64 matmul."""
import pytest
import torch

from helpers import rel_fro

pytestmark = pytest.mark.gpu


def _selftest(A, R, mode, use_rna):
    from banet_b200 import _lib
    lib = _lib.load(); _lib.require_device()
    D = torch.full((128, 160), float("nan"), device="cuda")
    _lib.check(lib.banet_tc_selftest(A.data_ptr(), R.data_ptr(), D.data_ptr(), mode, use_rna, torch.cuda.current_stream().cuda_stream),
               "banet_tc_selftest")
    torch.cuda.synchronize()
    return D


def _tf32_exact(x):
    return (x.view(torch.int32) & -8192).view(torch.float32)          # 10-bit mantissa: exactly representable in tf32


def test_tcgen05_layout_exact_on_tf32_representable_inputs():
    g = torch.Generator().manual_seed(0)
    A = _tf32_exact(torch.randn(64, 128, generator=g)).cuda(); R = _tf32_exact(torch.randn(64, 160, generator=g)).cuda()
    D = _selftest(A, R, 0, 0)
    ref = A.double().t() @ R.double()
    assert torch.isfinite(D).all()
    assert rel_fro(D, ref) < 1e-6          # only fp32 accumulation error remains: layout + descriptors are right


def test_tcgen05_index_pattern():
    """A and R with one-hot structure: D[i,j] must pick exactly A[k,i]*R[k,j] — catches any swizzle/stride mix-up."""
    A = torch.zeros(64, 128); R = torch.zeros(64, 160)
    for k in range(64):
        A[k, (3 * k + 1) % 128] = float(k + 1); R[k, (7 * k + 2) % 160] = 1.0
    D = _selftest(A.cuda(), R.cuda(), 0, 0)
    assert torch.equal(D.cpu(), A.t() @ R)


@pytest.mark.parametrize("mode,use_rna,tol", [(0, 0, 2e-3), (0, 1, 1e-3), (1, 1, 3e-4), (1, 0, 1e-3)])
def test_tcgen05_precision_modes(mode, use_rna, tol):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(64, 128, generator=g).cuda(); R = torch.randn(64, 160, generator=g).cuda()
    D = _selftest(A, R, mode, use_rna)
    err = rel_fro(D, A.double().t() @ R.double())
    print(f"mode={mode} rna={use_rna} rel-fro={err:.3e}")
    assert err < tol
