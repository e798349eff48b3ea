"""GPU: bit-reproducibility.  The library promises results that are a pure function of the inputs (fixed-order reductions, no atomics on the
forward path, stochastic tf32 rounding hashed from the iterate): the same call twice must agree bit for bit, whatever the caller-owned
workspace held before (it is handed over uninitialised), and AUTO must equal the level-wise policy it resolves to."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _poisoned_ws(pattern):
    def make(nbytes, device):
        n = max(int(nbytes), 256)
        if pattern == "nan":
            return torch.full((n,), 0xFF, dtype=torch.uint8, device=device)          # every float a NaN
        g = torch.Generator(device="cuda").manual_seed(n % 9973 + 1)
        return torch.randint(0, 256, (n,), dtype=torch.uint8, device=device, generator=g)
    return make


@pytest.fixture(scope="module")
def scene():
    from banet_b200 import ops, synth
    from helpers import O
    sc = synth.make_scene(nb=3, H=240, W=320, C=128, K=128, level_ids=(2, 3), seed=77, device="cuda", dtype=torch.float32)
    packed = [ops.pack_mlp(O.init_lambda_mlp(128, seed=7 + l.level, dtype=torch.float32)).cuda() for l in sc.levels]
    levels = [ops.Level(l.conv1, l.conv2, l.intr, l.p, l.D, l.B, grid=l.grid) for l in sc.levels]      # 160x120 (TF32X3 under AUTO), 320x240 (TF32X1)
    return sc, levels, packed


def test_whole_solves_are_bit_reproducible(scene, monkeypatch):
    from banet_b200 import ops, _lib
    sc, levels, packed = scene
    for prec in (_lib.PREC_AUTO, _lib.PREC_FP32_SIMT, _lib.PREC_TF32X1, _lib.PREC_TF32X2, _lib.PREC_TF32X3):
        outs = []
        for pattern in ("nan", "random", "random"):
            monkeypatch.setattr(ops, "_ws", _poisoned_ws(pattern))
            outs.append(ops.lm_run(levels, 3, sc.R0, sc.T0, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0, precision=prec))
        for o in outs[1:]:
            for a, b, name in zip(outs[0], o, "RTWs"):
                assert torch.equal(a, b), (prec, name, float((a.double() - b.double()).abs().max()))
        assert int(outs[0][3].abs().max()) == 0
    monkeypatch.setattr(ops, "_ws", _poisoned_ws("nan"))
    a = ops.lm_run(levels, 3, sc.R0, sc.T0, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0, precision=_lib.PREC_AUTO)
    b = ops.lm_run(levels, 3, sc.R0, sc.T0, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0, precision=_lib.PREC_TF32_LEVELWISE)
    for x, y in zip(a, b):
        assert torch.equal(x, y)                 # AUTO is the level-wise policy at K = 128


@pytest.mark.parametrize("layout", ["3c", "f2"])
def test_builds_and_steps_are_bit_reproducible(scene, monkeypatch, layout):
    from banet_b200 import ops, _lib
    sc, levels, packed = scene
    for li, lv in enumerate(levels):
        L = lv if layout == "3c" else ops.Level(lv.conv1, lv.conv2[..., :128].contiguous(), lv.intr, lv.p, lv.D, lv.B, grid=lv.grid)
        for prec in (_lib.PREC_FP32_SIMT, _lib.PREC_TF32X1, _lib.PREC_TF32X2, _lib.PREC_TF32X3):
            outs = []
            for pattern in ("nan", "random"):
                monkeypatch.setattr(ops, "_ws", _poisoned_ws(pattern))
                outs.append(ops.lm_build(L, sc.R0, sc.T0, sc.W0, precision=prec))
            for a, b in zip(*outs):
                assert torch.equal(a, b), (layout, li, prec)
            assert bool(torch.isfinite(outs[0][0]).all())
        H, g, rbar, _nv = outs[0]
        s1 = ops.lm_step(H, g, rbar, lv.conv1.shape[1], packed[li], 1000.0, sc.R0, sc.T0, sc.W0)
        s2 = ops.lm_step(H, g, rbar, lv.conv1.shape[1], packed[li], 1000.0, sc.R0, sc.T0, sc.W0)
        for a, b in zip(s1, s2):
            assert torch.equal(a, b), (layout, li, "lm_step")
        n_pts = lv.conv1.shape[1]
        assert torch.equal(ops.lm_lambda(rbar, n_pts, packed[li], 1000.0), ops.lm_lambda(rbar, n_pts, packed[li], 1000.0))


def test_captured_graph_replays_the_eager_solve_bit_for_bit(scene):
    """ops.LMRunGraph: banet_lm_run captured once into a CUDA graph (nothing in the call allocates or synchronises)."""
    from banet_b200 import ops, _lib
    sc, levels, packed = scene
    eager = ops.lm_run(levels, 3, sc.R0, sc.T0, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0)
    gr = ops.LMRunGraph(levels, 3, mlp_packed=packed, l2_regularizer_base=1000.0)
    for _ in range(2):                                    # replays do not depend on what the previous replay left in the static buffers
        out = gr.solve(sc.R0, sc.T0, sc.W0)
        for a, b in zip(eager, out):
            assert torch.equal(a, b)
    R2 = sc.R0.clone(); T2 = sc.T0 * 0.5
    for a, b in zip(ops.lm_run(levels, 3, R2, T2, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0), gr.solve(R2, T2, sc.W0)):
        assert torch.equal(a, b)
