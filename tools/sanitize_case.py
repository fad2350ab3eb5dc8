"""One small pass through every C-ABI entry point, for compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool racecheck python tools/sanitize_case.py [N]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fastbox_b200 import _lib  # noqa: E402
from fastbox_b200 import kspace as ks  # noqa: E402
from _util import pk_function, transfer_fn  # noqa: E402

F = _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = 500.0
rng = np.random.default_rng(1)
plan = _lib.Plan(N, L, L, L)
_, pkf = pk_function(0.8)
n3 = N ** 3
for exact_below in (512, 8):                             # integer LUT, then the float-bit table (fast prologue path)
    with np.errstate(all="ignore"):
        mode, tab, l0, dl = ks.choose_sqrt_pk_table(pkf, N, L, L, L, N ** 6. / L ** 3, exact_below=exact_below)
    plan.set_sqrt_pk(tab, mode, l0, dl)
    ft = ks.filter_tables(transfer_fn, N, L, L, L)
    plan.set_filter(ft.tperp, ft.tpar, ft.tdense)
    plan.set_pk_bins(ks.bin_thresholds(ks.pk_bin_edges(2 * np.pi / L, 2 * np.pi * np.sqrt(3.) * N / L, 20)))
    re = plan.upload(rng.standard_normal(n3).astype(np.float32))
    im = plan.upload(rng.standard_normal(n3).astype(np.float32))
    field, f2, f3 = plan.alloc(n3 * 4), plan.alloc(n3 * 4), plan.alloc(n3 * 4)
    spec = plan.alloc((N // 2 + 1) * N * N * 8)
    flags = F.F_SQRTPK | F.F_FILTER
    plan.realise(re, im, flags=flags, field_out=field, spec_out=spec, want_pk=True)
    plan.realise(None, None, seed=3, flags=flags, field_out=field, want_pk=True)
    plan.spectrum_to_field(spec, f2, flags=F.F_EXP, scale=0.8)
    plan.spectrum_to_field(spec, f3, kind=F.KIND_VEL_Z, scale=100.0)
    plan.field_to_spectrum(field, want_pk=True)
    plan.field_to_spectrum(field, want_pk=True, poles=True)
    plan.pk_from_spectrum(spec)
    print("sqrt(P) mode", mode, "ok", flush=True)
out = plan.alloc(n3 * 4)
plan.rsd_remap(f2, f3, None, np.linspace(-0.5 * L, 0.5 * L, N), 100.0, out)
x = np.arange(N) - N / 2.
beam = plan.upload(np.exp(-0.5 * (x[:, None, None] ** 2 + x[None, :, None] ** 2) / (2.0 + 0 * x[None, None, :]) ** 2)
                   .astype(np.float32))
plan.beam_convolve(beam, field, out)
u = plan.upload(rng.random(n3))
counts = plan.alloc(n3 * 4)
plan.halo_counts(field, np.array([0.5], np.float32), 0, np.array([1.0], np.float32), 0, False, 0.0, u, counts)
nh = plan.halo_catalogue(counts)
cat = plan.alloc(max(nh, 1) * 24)
plan.halo_catalogue(counts, None, cat, nh)
plan.fg_cube(rng.uniform(1, 2, (N, N)), rng.normal(-2.5, 0.1, (N, N)), np.log2(np.linspace(700, 1100, N) / 130.), out)
plan.radiometer_noise(np.linspace(0.5, 1.5, N), out, None, seed=2, accumulate=True)
plan.sync()
print("all entry points ok; halos", nh)
