"""Forward + backward of one differentiable LM iteration on the fused kernels (banet_b200.autograd.iteration_fused), timed with CUDA events.
Shapes: the reference's training scale (nb=2, 4096 sampled points -> here a 64x64 grid) and dense levels of the cfg2 workload."""
import json, os, sys, statistics, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from banet_b200 import synth, autograd as ag, _lib
_lib.require_device()
rows = []
for nb, gh, gw in ((2, 64, 64), (8, 120, 160), (8, 240, 320), (8, 480, 640)):
    C = K = 128
    sc = synth.make_scene(nb=nb, H=gh, W=gw, C=C, K=K, level_ids=(3,), seed=21, device="cuda", dtype=torch.float32)
    lv = sc.levels[0]
    g = torch.Generator().manual_seed(7); dims = [C, 2 * C, 4 * C, 2 * C, C, 1]       # he_normal filters, zero biases (bundlenet.py:102-110)
    mlp = [((torch.randn(dims[i], dims[i + 1], generator=g) * (2.0 / dims[i]) ** 0.5).cuda().requires_grad_(), torch.zeros(dims[i + 1], device="cuda").requires_grad_())
           for i in range(5)]
    leaf = lambda t: t.detach().clone().requires_grad_()
    conv1, conv2, D, B, R, T, W = leaf(lv.conv1), leaf(lv.conv2), leaf(lv.D), leaf(lv.B), leaf(sc.R0), leaf(sc.T0), leaf(sc.W0)
    lam = None if mlp else torch.full((nb,), 0.5, device="cuda")
    def fwd():
        return ag.iteration_fused(conv1, conv2, lv.intr, lv.p, D, B, R, T, W, mlp, 1000.0 if mlp else None, lambda_override=lam, grid=lv.grid)
    def ev(): return torch.cuda.Event(enable_timing=True)
    tf, tb = [], []
    for it in range(6):
        for t in (conv1, conv2, D, B, R, T, W): t.grad = None
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(); Rn, Tn, Wn = fwd(); loss = Rn.sum() + Tn.sum() + (Wn * Wn).sum(); e1.record(); loss.backward(); e2.record()
        torch.cuda.synchronize()
        if it >= 2: tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
    N = gh * gw
    row = {"nb": nb, "level": f"{gw}x{gh}", "N": N, "C": C, "K": K, "lambda": "mlp" if mlp else "fixed", "forward_ms": statistics.median(tf), "backward_ms": statistics.median(tb),
           "saved_for_backward": "inputs + H, g, delta only (J, G, d never materialised)",
           "reference_graph_would_materialise_GB": nb * N * (2 * (K + 6) + 3 * C) * 4 / 1e9}
    rows.append(row); print(row, flush=True)
    del conv1, conv2, D, B, sc, lv
    torch.cuda.empty_cache()
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
json.dump(rows, open(os.path.join(out, "training_step_timing.json"), "w"), indent=1)
