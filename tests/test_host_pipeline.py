"""Host-buffer pipeline: chunking logic on CPU; chunked, overlapped solve equals the plain device solve on the GPU."""
import pytest
import torch

from banet_b200.host_pipeline import chunk_ranges


def test_chunk_ranges_cover_and_balance():
    for nb in (1, 2, 5, 32, 33):
        for ch in (1, 2, 4, 7, 64):
            r = chunk_ranges(nb, ch)
            assert r[0][0] == 0 and r[-1][1] == nb
            assert all(a < b for a, b in r) and all(r[i][1] == r[i + 1][0] for i in range(len(r) - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and len(r) == min(ch, nb)


@pytest.mark.gpu
@pytest.mark.parametrize("chunks,derive", [(1, True), (3, True), (2, False)])
def test_host_solver_matches_device_solve(chunks, derive):
    from banet_b200 import ops, synth
    from banet_b200.host_pipeline import HostSolver
    from helpers import rel_fro
    nb, C, K = 5, 64, 128
    sc = synth.make_scene(nb=nb, H=96, W=128, C=C, K=K, level_ids=(2, 3), seed=23, device="cuda", dtype=torch.float32)
    g = torch.Generator().manual_seed(7)
    dims = [C, 2 * C, 4 * C, 2 * C, C, 1]
    packed = []
    for _ in sc.levels:
        params = [(torch.randn(dims[i], dims[i + 1], generator=g) * (2.0 / dims[i]) ** 0.5, torch.zeros(dims[i + 1])) for i in range(5)]
        packed.append(ops.pack_mlp(params).cuda())
    host = []
    for l in sc.levels:
        conv2 = l.conv2[..., :C].contiguous() if derive else l.conv2
        host.append({"conv1": l.conv1.cpu().pin_memory(), "conv2": conv2.cpu().pin_memory(), "intr": l.intr.cpu().pin_memory(),
                     "p": l.p.cpu().pin_memory(), "D": l.D.cpu().pin_memory(), "B": l.B.cpu().pin_memory(), "grid": l.grid})
    hs = HostSolver(host, derive_gradients=derive, chunks=chunks, precision=0)
    oR = torch.empty(nb, 3, 3).pin_memory(); oT = torch.empty(nb, 3, 1).pin_memory(); oW = torch.empty(nb, K, 1).pin_memory()
    R, T, W, st = hs.solve(sc.R0.cpu().pin_memory(), sc.T0.cpu().pin_memory(), sc.W0.cpu().pin_memory(), 3, mlp_packed=packed,
                           l2_regularizer_base=1000.0, out=(oR, oT, oW))
    torch.cuda.synchronize()
    levels = [ops.Level(l.conv1, ops.grad_fixed_concat(l.conv2[..., :C].contiguous()) if derive else l.conv2, l.intr, l.p, l.D, l.B, grid=l.grid)
              for l in sc.levels]
    R1, T1, W1, st1 = ops.lm_run(levels, 3, sc.R0, sc.T0, sc.W0, mlp_packed=packed, l2_regularizer_base=1000.0, precision=0)
    assert int(st.abs().max()) == 0 and int(st1.abs().max()) == 0
    # a different batch size moves the CTA / partial-slot boundaries: same maths, different fp32 summation order
    assert rel_fro(R, R1) < 1e-6 and rel_fro(T, T1) < 1e-5 and rel_fro(W, W1) < 1e-4
    assert torch.equal(oR, R.cpu()) and torch.equal(oT, T.cpu()) and torch.equal(oW, W.cpu())


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", [2, 4])
def test_resize_host_solver_matches_explicit_derivation(chunks):
    """The BundleResize-boundary host pipeline (zero-copy conv1, other-half conv2, D/B/p derived on the device, chunked + overlapped copies)
    against the same solve with every level tensor derived explicitly through the tested ops, and against the planted solution."""
    from banet_b200 import ops, synth
    from banet_b200.host_pipeline import ResizeHostSolver
    from helpers import rel_fro
    nimg, C, K = 8, 64, 128
    sc = synth.make_resize_scene(nimg, 96, 128, C, K, level_ids=(2, 3), seed=41, device="cuda")
    pin = lambda t: t.cpu().pin_memory()
    hs = ResizeHostSolver([pin(l) for l in sc.layers], pin(sc.basis), pin(sc.init_depth), pin(sc.intr), sc.scales, chunks=chunks, precision=0)
    R, T, W, st = hs.solve(pin(sc.R0), pin(sc.T0), pin(sc.W0), 4, lambda_fixed=0.5)
    torch.cuda.synchronize()
    assert int(st.abs().max()) == 0
    half = nimg // 2
    levels = []
    for lay, s in zip(sc.layers, sc.scales):
        h, w = lay.shape[1], lay.shape[2]
        vv, uu = torch.meshgrid(torch.arange(h, device="cuda", dtype=torch.float32), torch.arange(w, device="cuda", dtype=torch.float32), indexing="ij")
        pts = torch.stack([uu.reshape(-1), vv.reshape(-1)], -1).unsqueeze(0).repeat(nimg, 1, 1).contiguous()
        conv1 = ops.resample(lay, pts, 1.0)                                   # reference: layer1 = resampler(layers[level], points / scale)
        conv2 = torch.cat([lay[half:], lay[:half]], 0).contiguous()           # reference: the half swap
        intr_l = sc.intr / s
        levels.append(ops.Level(conv1, conv2, intr_l, ops.compute_coordinates(pts, intr_l, True), ops.resample(sc.init_depth, pts, s / 2.0),
                                ops.resample(sc.basis, pts, s / 2.0), grid=(w, h)))
    R1, T1, W1, st1 = ops.lm_run(levels, 4, sc.R0, sc.T0, sc.W0, lambda_fixed=0.5, precision=0)
    # a different batch size per call moves the CTA / partial-slot boundaries: same maths, different fp32 summation order
    assert rel_fro(R, R1) < 2e-5 and rel_fro(T, T1) < 1e-3 and rel_fro(W, W1) < 5e-3


def test_numa_binding_follows_sysfs_and_restores_the_affinity(tmp_path):
    """host_pipeline.numa_local_to: the calling thread runs on the CPUs of the GPU's NUMA node while host buffers are allocated, and gets its
    affinity back afterwards; anything unreadable makes it a no-op.  (A fake sysfs tree: no GPU needed.)"""
    import os
    from banet_b200.host_pipeline import numa_local_to
    before = os.sched_getaffinity(0)
    cpus = sorted(before)
    node_cpus = cpus[: max(1, len(cpus) // 2)]
    bdf = "0000:1b:00.0"
    (tmp_path / "bus/pci/devices" / bdf).mkdir(parents=True)
    (tmp_path / "bus/pci/devices" / bdf / "numa_node").write_text("1\n")
    (tmp_path / "devices/system/node/node1").mkdir(parents=True)
    lo, hi = node_cpus[0], node_cpus[-1]
    listed = [c for c in range(lo, hi + 1)]                                  # a range like the kernel prints it, plus a CPU we may not own
    (tmp_path / "devices/system/node/node1/cpulist").write_text(f"{lo}-{hi},{max(cpus) + 1000}\n")
    with numa_local_to("cuda:0", _bdf=bdf, _sysfs=str(tmp_path)) as n:
        inside = os.sched_getaffinity(0)
    assert inside == set(listed) & before and n.info["node"] == 1 and n.info["cpus"] == len(inside)
    assert os.sched_getaffinity(0) == before
    (tmp_path / "bus/pci/devices" / bdf / "numa_node").write_text("-1\n")          # unknown node: nothing changes
    with numa_local_to("cuda:0", _bdf=bdf, _sysfs=str(tmp_path)) as n:
        assert os.sched_getaffinity(0) == before
    assert n.info["node"] is None and os.sched_getaffinity(0) == before
    with numa_local_to("cuda:0", _bdf="ffff:ff:1f.0", _sysfs=str(tmp_path)) as n:   # unreadable topology: a no-op, never an exception
        assert os.sched_getaffinity(0) == before
    assert "unavailable" in n.info["note"]
