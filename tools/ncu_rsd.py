import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastbox_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L = 2000.0 * N / 1024
plan = _lib.Plan(N, L, L, L)
rng = np.random.default_rng(0)
d = plan.upload(rng.standard_normal(N ** 3).astype(np.float32))
v = plan.upload((300.0 * rng.standard_normal(N ** 3)).astype(np.float32))
out = plan.alloc(N ** 3 * 4)
z = np.linspace(-0.5 * L, 0.5 * L, N)
for it in range(3):
    plan.timer_start()
    plan.rsd_remap(d, v, None, z, 100.0, out)
    print("rsd ms", plan.timer_stop())
u = plan.upload(rng.random(N ** 3))
counts = plan.alloc(N ** 3 * 4)
for it in range(2):
    plan.timer_start()
    plan.halo_counts(d, np.array([1e-3], np.float32), 0, np.array([1.0], np.float32), 0, False, 0.0, u, counts)
    print("halo ms", plan.timer_stop())
