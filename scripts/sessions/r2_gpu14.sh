#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/g14_*
timeout -s KILL 300 python -m pytest "tests/test_gpu_parity.py" -q -m gpu --timeout 120 > gpurun_out/g14_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g14_rc.txt
cat > /tmp/step.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from banet_b200 import ops
g = torch.Generator().manual_seed(1)
nb, K, C, N = 32, 128, 128, 307200
P = 6 + K
A = torch.randn(nb, P, 3 * P, generator=g, dtype=torch.float64)
H = ((A @ A.transpose(1, 2)) / (3 * P)).float().cuda(); gv = (torch.randn(nb, P, generator=g) * 1e-2).cuda()
rbar = (torch.rand(nb, C, generator=g) * N * 0.2).cuda()
dims = [C, 2 * C, 4 * C, 2 * C, C, 1]
mlp = ops.pack_mlp([(torch.randn(dims[i], dims[i + 1], generator=g) * (2.0 / dims[i]) ** 0.5, torch.zeros(dims[i + 1])) for i in range(5)]).cuda()
R = torch.eye(3).repeat(nb, 1, 1).cuda(); T = torch.zeros(nb, 3, 1).cuda(); W = torch.zeros(nb, K, 1).cuda()
for _ in range(3): ops.lm_step(H, gv, rbar, N, mlp, 1000.0, R, T, W)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.lm_step(H, gv, rbar, N, mlp, 1000.0, R, T, W)
e1.record(); torch.cuda.synchronize()
print("lm_step us (incl. python call):", e0.elapsed_time(e1) / 20 * 1000)
PY
timeout -s KILL 120 python /tmp/step.py > gpurun_out/g14_step_time.log 2>&1
cat gpurun_out/g14_rc.txt; tail -2 gpurun_out/g14_tests.log; cat gpurun_out/g14_step_time.log
