"""Does the power-of-two plane stride hurt the x passes?  Time them with ncols = N^2 + pad (pad columns of slack)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastbox_b200 import _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
plan = _lib.Plan(N, 2000., 2000., 2000.)
for pad in (0, 16, 32, 64, 256, 1024, 4096, 16 * 1031):
    ncols = N * N + pad
    spec = plan.alloc((N // 2 + 1) * ncols * 8)
    field = plan.alloc(N * ncols * 4)
    plan.affine(field, N * ncols, 0.0, 1.0)
    plan.fft_pass_x_r2c(field, spec, ncols)
    plan.fft_pass_x_c2r(spec, field, ncols, scale=1.0 / N)
    plan.sync()
    t_r2c, t_c2r = [], []
    for _ in range(3):
        plan.timer_start()
        plan.fft_pass_x_r2c(field, spec, ncols)
        t_r2c.append(plan.timer_stop())
        plan.timer_start()
        plan.fft_pass_x_c2r(spec, field, ncols, scale=1.0 / N)
        t_c2r.append(plan.timer_stop())
    print("pad %6d  ncols %9d  r2c %.3f ms  c2r %.3f ms" % (pad, ncols, min(t_r2c), min(t_c2r)), flush=True)
    spec.free()
    field.free()
