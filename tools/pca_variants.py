"""PCA covariance (fb_pca_covariance, filters.py:158-159) at N^3 float64: kernel variants selected by environment.

    python tools/pca_variants.py [N] [variant ...]      variant = tile:kt:waves, e.g. 64:0:0 128:8:9 128:16:9
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastbox_b200 import _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
variants = sys.argv[2:] or ["64:0:0", "128:8:9", "128:16:9", "128:8:18", "128:8:4"]
plan = _lib.Plan(N, 1.0, 1.0, 1.0)
rng = np.random.default_rng(0)
f32 = plan.upload(rng.standard_normal(N ** 3).astype(np.float32))
cube = plan.alloc(N ** 3 * 8)
plan.lib.fb_convert_f32_to_f64(plan.h, _lib._ptr(f32), _lib._ptr(cube), N ** 3)
flops = 2.0 * (N * (N + 1) / 2.0) * N * N           # the upper triangle, nothing redundant
ref = None
for v in variants:
    tile, kt, waves = (int(t) for t in v.split(":"))
    os.environ["FB_PCA_TILE"] = str(tile)
    if kt:
        os.environ["FB_PCA_KT"] = str(kt)
    if waves:
        os.environ["FB_PCA_WAVES"] = str(waves)
    best = 1e30
    for it in range(int(os.environ.get("PCA_REPS", "3"))):
        plan.timer_start()
        mean, cov = plan.pca_covariance(cube)
        ms = plan.timer_stop()
        if it or os.environ.get("PCA_REPS") == "1":
            best = min(best, ms)
    if ref is None:
        ref = cov
    dev = np.max(np.abs(cov - ref)) / np.max(np.abs(ref))
    print("pca cov N=%d tile=%d kt=%d waves=%d: %.2f ms  %.1f TFLOP/s  (max dev from first variant %.1e)"
          % (N, tile, kt, waves, best, flops / (best * 1e-3) / 1e12, dev), flush=True)
