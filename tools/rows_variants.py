"""Time the passes under different flag combinations to see what costs what."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastbox_b200 import _lib  # noqa: E402
from fastbox_b200 import kspace as ks  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
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
re = plan.alloc(N ** 3 * 4)
im = plan.alloc(N ** 3 * 4)
spec = plan.alloc(h * N * N * 8)
F = _lib
for label, src, flags, pk in [("noise plain", 1, 0, False), ("noise sqrtpk", 1, F.F_SQRTPK, False),
                              ("noise sqrtpk+filter", 1, F.F_SQRTPK | F.F_FILTER, False),
                              ("noise sqrtpk+pk", 1, F.F_SQRTPK, True),
                              ("noise all", 1, F.F_SQRTPK | F.F_FILTER, True),
                              ("philox plain", 0, 0, False), ("philox all", 0, F.F_SQRTPK | F.F_FILTER, True)]:
    best = None
    for it in range(3):
        plan.realise(re if src else None, im if src else None, seed=it, flags=flags, field_out=field, want_pk=pk)
        t = plan.last_timings(3)
        best = t if best is None or t[0] < best[0] else best
    print("%-22s rows %.3f cols %.3f x %.3f ms" % (label, best[0], best[1], best[2]), flush=True)
plan.realise(re, im, flags=F.F_SQRTPK, field_out=field, spec_out=spec)
for label, flags in [("spec plain", 0), ("spec filter", F.F_FILTER)]:
    for it in range(3):
        plan.spectrum_to_field(spec, field, flags=flags)
        t = plan.last_timings(3)
    print("%-22s rows %.3f ms" % (label, t[0]), flush=True)
for pk in (False, True):
    for it in range(3):
        plan.field_to_spectrum(field, want_pk=pk)
        t = plan.last_timings(3)
    print("forward pk=%s: x %.3f cols %.3f rows %.3f" % (pk, t[0], t[1], t[2]), flush=True)
