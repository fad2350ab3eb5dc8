"""One short realise(+filter+P(k)) + forward P(k) at size N for ncu captures."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastbox_b200 import _lib  # noqa: E402
from fastbox_b200 import kspace as ks  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
src_mode = sys.argv[2] if len(sys.argv) > 2 else "philox"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
L = 2000.0
plan = _lib.Plan(N, L, L, L)
pkf = lambda k: np.where(k > 0, 1e4 * (k / 0.02) / (1 + (k / 0.02) ** 2.5), 0.0)
with np.errstate(all="ignore"):
    mode, tab, l0, dl = ks.choose_sqrt_pk_table(pkf, N, L, L, L, N ** 6. / L ** 3,
                                                exact_below=int(os.environ.get("FB_EXACT_BELOW", "512")))
plan.set_sqrt_pk(tab, mode, l0, dl)
plan.set_pk_bins(ks.bin_thresholds(ks.pk_bin_edges(2 * np.pi / L, 2 * np.pi * np.sqrt(3.) * N / L, 50)))
m = ks.mode_numbers(N).astype(np.float64)
h = N // 2 + 1
kperp = 2 * np.pi * np.sqrt((m[:h, None] / L) ** 2 + (m[None, :] / L) ** 2)
kpar = 2 * np.pi * m / L
plan.set_filter(np.exp(-0.5 * (kperp / 0.1) ** 2), 1. - np.exp(-0.5 * (kpar / 0.001) ** 2), None)
field = plan.alloc(N ** 3 * 4)
re = im = None
if src_mode == "noise":
    re = plan.alloc(N ** 3 * 4)
    im = plan.alloc(N ** 3 * 4)
for it in range(reps):
    plan.realise(re, im, seed=it, flags=_lib.F_SQRTPK | _lib.F_FILTER, field_out=field, want_pk=True)
    print("realise", plan.last_timings(3))
    plan.field_to_spectrum(field, want_pk=True)
    print("forward", plan.last_timings(3))
