"""Solve frame-pairs whose level tensors live in host memory: the call a user makes when the feature pyramid comes off a
data loader / another process.  Mirrors the layer boundary of the reference's `BundleResize` (bundlenet.py:376-399): feature
maps in, `[F2|gx|gy]` derived on the device (:386-389), coarse-to-fine LM solve (:376-399), (R, T, W) out.

Pairs are independent, so the batch is cut into chunks: the host->device copies of chunk k+1 (copy stream, pinned buffers)
overlap the solve of chunk k (compute stream); the PCIe transfer is the long pole (20 GB per 32-pair cfg2 batch vs 51 ms of
compute), so the solve hides behind it except for the last chunk.  No CPU fallback: everything after the copy is the C-ABI."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import PREC_AUTO

Tensor = torch.Tensor
_NAMES = ("conv1", "conv2", "intr", "p", "D", "B")


def chunk_ranges(nb: int, chunks: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced pair ranges (first ranges one longer); never empty ranges."""
    chunks = max(1, min(int(chunks), int(nb)))
    base, rem = divmod(nb, chunks)
    out, a = [], 0
    for c in range(chunks):
        b = a + base + (1 if c < rem else 0)
        out.append((a, b)); a = b
    return out


class HostSolver:
    """host_levels: one dict per level (coarse -> fine) with host tensors conv1 [nb,N,C], conv2 [nb,h,w,C] (features; gradients
    are derived on the device) or [nb,h,w,3C] (already [F2|gx|gy]) when derive_gradients=False, intr [nb,4], p [nb,3,N],
    D [nb,N,1], B [nb,N,K], and optionally grid=(w,h).  Pinned host tensors make the copies asynchronous."""

    def __init__(self, host_levels: Sequence[Dict], derive_gradients: bool = True, chunks: int = 4, device=None,
                 precision: int = PREC_AUTO):
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.host = list(host_levels)
        self.derive = bool(derive_gradients)
        self.precision = precision
        self.nb = int(self.host[0]["conv1"].shape[0])
        self.ranges = chunk_ranges(self.nb, chunks)
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.compute_stream = torch.cuda.Stream(self.dev)
        self.dev_levels: List[Dict[str, Tensor]] = []
        self.h2d_bytes = 0
        for hl in self.host:
            d = {}
            for name in _NAMES:
                t = hl[name]
                d[name] = torch.empty(t.shape, dtype=torch.float32, device=self.dev)
                self.h2d_bytes += t.numel() * 4
            if self.derive:                       # staging of the feature maps; conv2 on the device is the 3C tensor
                C = hl["conv2"].shape[-1]
                d["feat"] = d["conv2"]
                d["conv2"] = torch.empty(*hl["conv2"].shape[:-1], 3 * C, dtype=torch.float32, device=self.dev)
            self.dev_levels.append(d)
        self._ws: Optional[Tensor] = None

    def _levels(self, a: int, b: int) -> List[ops.Level]:
        return [ops.Level(d["conv1"][a:b], d["conv2"][a:b], d["intr"][a:b], d["p"][a:b], d["D"][a:b], d["B"][a:b], grid=hl.get("grid"))
                for d, hl in zip(self.dev_levels, self.host)]

    def solve(self, R0: Tensor, T0: Tensor, W0: Tensor, iters_per_level: int, mlp_packed=None, l2_regularizer_base: float = 1000.0,
              lambda_fixed: float = -1.0, out: Optional[Tuple[Tensor, Tensor, Tensor]] = None):
        """R0 [nb,3,3], T0 [nb,3,1], W0 [nb,K,1] on the host (pinned for async copies).  Returns device (R, T, W, status); when
        `out` = three host tensors is given they receive the results too (device->host inside the pipeline)."""
        nb, dev = self.nb, self.dev
        R = torch.empty(nb, 3, 3, device=dev); T = torch.empty(nb, 3, 1, device=dev); W = torch.empty(W0.shape, device=dev)
        status = torch.empty(nb, dtype=torch.int32, device=dev)
        dR = torch.empty_like(R); dT = torch.empty_like(T); dW = torch.empty_like(W)
        self.copy_stream.wait_stream(torch.cuda.current_stream(dev))
        self.compute_stream.wait_stream(torch.cuda.current_stream(dev))
        events = []
        with torch.cuda.stream(self.copy_stream):
            for a, b in self.ranges:
                for d, hl in zip(self.dev_levels, self.host):
                    for name in _NAMES:
                        dst = d["feat"] if (name == "conv2" and self.derive) else d[name]
                        dst[a:b].copy_(hl[name][a:b], non_blocking=True)
                dR[a:b].copy_(R0[a:b], non_blocking=True); dT[a:b].copy_(T0[a:b], non_blocking=True); dW[a:b].copy_(W0[a:b], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(self.copy_stream); events.append(ev)
        with torch.cuda.stream(self.compute_stream):
            for (a, b), ev in zip(self.ranges, events):
                self.compute_stream.wait_event(ev)
                if self.derive:
                    for d in self.dev_levels:
                        ops.grad_fixed_concat(d["feat"][a:b], out=d["conv2"][a:b])
                lv = self._levels(a, b)
                if self._ws is None:
                    self._ws = torch.empty(ops.lm_run_workspace_bytes(lv, self.precision), dtype=torch.uint8, device=dev)
                r, t, w, st = ops.lm_run(lv, iters_per_level, dR[a:b], dT[a:b], dW[a:b], mlp_packed=mlp_packed,
                                         l2_regularizer_base=l2_regularizer_base, lambda_fixed=lambda_fixed, workspace=self._ws,
                                         precision=self.precision)
                R[a:b] = r; T[a:b] = t; W[a:b] = w; status[a:b] = st
                if out is not None:
                    out[0][a:b].copy_(r, non_blocking=True); out[1][a:b].copy_(t, non_blocking=True); out[2][a:b].copy_(w, non_blocking=True)
        torch.cuda.current_stream(dev).wait_stream(self.compute_stream)
        for t_ in (R, T, W, status, dR, dT, dW):
            t_.record_stream(self.compute_stream)
        return R, T, W, status
