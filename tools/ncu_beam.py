"""Beam convolution + halo counts at size N for ncu captures: python tools/ncu_beam.py [N]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastbox_b200 import _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = 2000.0
plan = _lib.Plan(N, L, L, L)
n3 = N ** 3
field = plan.alloc(n3 * 4)
beam = plan.alloc(n3 * 4)
out = plan.alloc(n3 * 4)
plan.affine(field, n3, 0.0, 1.0)
plan.affine(beam, n3, 0.0, 1.0 / N ** 2)
plan.timer_start()
plan.beam_set(beam)
print("beam set-up ms", plan.timer_stop())
for _ in range(reps):
    plan.timer_start()
    plan.beam_convolve(None, field, out)
    print("beam ms (cached spectrum)", plan.timer_stop())
u = plan.upload(np.random.default_rng(1).random(n3)) if len(sys.argv) > 3 else plan.alloc(n3 * 8)
counts = plan.alloc(n3 * 4)
nbar = np.array([1e-3], np.float32)
bias = np.array([1.0], np.float32)
for _ in range(reps):
    plan.timer_start()
    plan.halo_counts(field, nbar, 0, bias, 0, False, 0.0, u, counts)
    print("halo ms", plan.timer_stop())
if len(sys.argv) > 3:                                    # catalogue of the counts just drawn
    plan.affine(field, n3, 0.0, 0.0)
    plan.halo_counts(field, np.array([float(sys.argv[3])], np.float32), 0, bias, 0, False, 0.0, u, counts)
    nh = plan.halo_catalogue(counts)
    cat = plan.alloc(max(nh, 1) * 24)
    for _ in range(reps):
        plan.timer_start()
        plan.halo_catalogue(counts, None, cat, nh)
        print("catalogue ms", plan.timer_stop(), "halos", nh)
