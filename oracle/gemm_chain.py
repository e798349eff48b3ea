"""Literal emulation of the reference's column-major cuBLAS call chain (TEST INFRASTRUCTURE).

`gemm_batched` reproduces cublasSgemmBatched semantics (column-major, leading dimensions, op(A)/op(B))
on flat row-major buffers exactly as utils.cu passes them, so the identities
    left  = sum_n J^T G^T G J        right = sum_n J^T G^T d           (utils.cu:331-414)
    dA = 2 A Ghat + d ghat^T, dd = A ghat, dJ = G^T dA, dG = dA J^T    (utils.cu:625-690)
claimed in oracle/ba_oracle.py are checked against the actual argument lists of the reference, not
against our reading of them.  numpy float64.
"""
from __future__ import annotations

import numpy as np


def _colmajor(buf: np.ndarray, rows: int, cols: int, ld: int) -> np.ndarray:
    """View flat `buf` as a column-major rows x cols matrix with leading dimension ld."""
    assert ld >= rows
    out = np.empty((rows, cols), dtype=buf.dtype)
    for c in range(cols):
        out[:, c] = buf[c * ld: c * ld + rows]
    return out


def _store_colmajor(buf: np.ndarray, M: np.ndarray, ld: int) -> None:
    rows, cols = M.shape
    for c in range(cols):
        buf[c * ld: c * ld + rows] = M[:, c]


def gemm(transa: bool, transb: bool, m: int, n: int, k: int, alpha: float,
         A: np.ndarray, lda: int, B: np.ndarray, ldb: int, beta: float, Cbuf: np.ndarray, ldc: int) -> None:
    """C = alpha * op(A) op(B) + beta * C, cuBLAS conventions, flat buffers."""
    a = _colmajor(A, k if transa else m, m if transa else k, lda)
    b = _colmajor(B, n if transb else k, k if transb else n, ldb)
    opa = a.T if transa else a
    opb = b.T if transb else b
    c = _colmajor(Cbuf, m, n, ldc) if beta != 0.0 else np.zeros((m, n), dtype=Cbuf.dtype)
    _store_colmajor(Cbuf, alpha * (opa @ opb) + beta * c, ldc)


def equation_construction_chain(J: np.ndarray, G: np.ndarray, d: np.ndarray):
    """utils.cu:331-414 for one batch entry.  J [N,2,P], G [N,C,2], d [N,C,1] (row-major, as TF passes them)."""
    N, jr, P = J.shape
    _, C, gc = G.shape
    dc = d.shape[2]
    left = np.zeros(P * P)
    right = np.zeros(P * dc)
    for n in range(N):
        Jn, Gn, dn = J[n].reshape(-1).copy(), G[n].reshape(-1).copy(), d[n].reshape(-1).copy()
        buf0 = np.zeros(gc * gc); buf1 = np.zeros(jr * P); buf2 = np.zeros(P * P)
        buf3 = np.zeros(gc * dc); buf4 = np.zeros(P * dc)
        gemm(False, True, gc, gc, C, 1.0, Gn, gc, Gn, gc, 0.0, buf0, gc)          # :331-340
        gemm(False, False, P, gc, jr, 1.0, Jn, P, buf0, gc, 0.0, buf1, P)         # :344-353
        gemm(False, True, P, P, gc, 1.0, buf1, P, Jn, P, 0.0, buf2, P)            # :356-365
        left += buf2                                                              # ColumnReduce :380
        gemm(False, True, dc, gc, C, 1.0, dn, dc, Gn, gc, 0.0, buf3, dc)          # :382-391
        gemm(False, True, dc, P, jr, 1.0, buf3, dc, Jn, P, 0.0, buf4, dc)         # :393-402
        right += buf4                                                             # :414
    return left.reshape(P, P), right.reshape(P, dc)


def equation_construction_grad_chain(J, G, d, left_grad, right_grad):
    """utils.cu:625-690 for one batch entry.  left_grad [P,P], right_grad [P,1] (row-major)."""
    N, jr, P = J.shape
    _, C, gc = G.shape
    dc = d.shape[2]
    dJ = np.zeros_like(J); dG = np.zeros_like(G); dd = np.zeros_like(d)
    g0 = left_grad.reshape(-1).copy()        # tile_kernel :613-617 copies these to every pixel
    g1 = right_grad.reshape(-1).copy()
    for n in range(N):
        Jn, Gn, dn = J[n].reshape(-1).copy(), G[n].reshape(-1).copy(), d[n].reshape(-1).copy()
        A = np.zeros(C * P); AG = np.zeros(C * P)
        o_dd = np.zeros(C * dc); o_dJ = np.zeros(jr * P); o_dG = np.zeros(C * gc)
        gemm(False, False, P, C, jr, 1.0, Jn, P, Gn, gc, 0.0, A, P)               # :625-634  A = G J
        gemm(False, False, dc, C, P, 1.0, g1, dc, A, P, 0.0, o_dd, dc)            # :636-645  dd = A ghat
        gemm(False, False, P, C, P, 2.0, g0, P, A, P, 0.0, AG, P)                 # :648-657  2 A Ghat
        gemm(True, False, P, C, dc, 1.0, g1, dc, dn, dc, 1.0, AG, P)              # :659-668  += d ghat^T
        gemm(False, True, P, gc, C, 1.0, AG, P, Gn, gc, 0.0, o_dJ, P)             # :670-679  dJ = G^T dA
        gemm(True, False, gc, C, P, 1.0, Jn, P, AG, P, 0.0, o_dG, gc)             # :681-690  dG = dA J^T
        dJ[n] = o_dJ.reshape(jr, P); dG[n] = o_dG.reshape(C, gc); dd[n] = o_dd.reshape(C, dc)
    return dJ, dG, dd
