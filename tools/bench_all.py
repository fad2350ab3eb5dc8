"""
Secondary benchmark: every hot-path row of SURVEY section 8 at one grid size, device-resident
buffers, CUDA events on the library stream, with the algorithmic bytes/cell of SURVEY 8(d).

    python tools/bench_all.py [N]         (default 512; 1024 needs ~40 GB)
"""
import json
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


def timed(plan, fn, reps=5):
    fn()
    plan.sync()
    best = 1e30
    for _ in range(reps):
        plan.timer_start()
        fn()
        best = min(best, plan.timer_stop())
    return best


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    L = 2000.0 * N / 1024
    peak = 6541.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    plan = _lib.Plan(N, L, L, L)
    _, pkf = pk_function(0.8)
    with np.errstate(all="ignore"):
        mode, tab, l0, dl = ks.choose_sqrt_pk_table(pkf, N, L, L, L, N ** 6. / L ** 3)
    plan.set_sqrt_pk(tab, mode, l0, dl)
    ft = ks.filter_tables(transfer_fn, N, L, L, L)
    plan.set_filter(ft.tperp, ft.tpar, ft.tdense)
    plan.set_pk_bins(ks.bin_thresholds(ks.pk_bin_edges(2 * np.pi / L, 2 * np.pi * np.sqrt(3.) * N / L, 50)))
    n3 = N ** 3
    nh = (N // 2 + 1) * N * N
    field = plan.alloc(n3 * 4)
    f2 = plan.alloc(n3 * 4)
    f3 = plan.alloc(n3 * 4)
    spec = plan.alloc(nh * 8)
    zgrid = np.linspace(-0.5 * L, 0.5 * L, N)
    rows = []

    def add(name, bytes_per_cell, ms):
        gbs = bytes_per_cell * n3 / (ms * 1e-3) / 1e9
        rows.append(dict(stage=name, ms=ms, Mcells_per_s=n3 / ms / 1e3, bytes_per_cell=bytes_per_cell, GBs=gbs,
                         frac_hbm=gbs / peak))
        print("%-46s %8.3f ms %10.0f Mcells/s %5.0f B/cell %7.0f GB/s %5.1f%%" %
              (name, ms, n3 / ms / 1e3, bytes_per_cell, gbs, 100 * gbs / peak), flush=True)

    add("realise (Philox) + store delta_k  [a3]", 24,
        timed(plan, lambda: plan.realise(None, None, seed=1, flags=F.F_SQRTPK, field_out=field, spec_out=spec)))
    add("realise+filter+P(k) (Philox)  [a3,a8,a10]", 20,
        timed(plan, lambda: plan.realise(None, None, seed=1, flags=F.F_SQRTPK | F.F_FILTER, field_out=f2,
                                         want_pk=True)))
    add("bias + exp (log-normal) from delta_k  [a4,a5]", 24,
        timed(plan, lambda: plan.spectrum_to_field(spec, f2, flags=F.F_EXP, scale=0.84)))
    add("log-normal normalise pass  [a4]", 8, timed(plan, lambda: plan.affine(f2, n3, 1.0, -1.0)))
    add("velocity v_z from delta_k  [a6]", 24,
        timed(plan, lambda: plan.spectrum_to_field(spec, f3, kind=F.KIND_VEL_Z, scale=100.0)))
    out = plan.alloc(n3 * 4)
    add("redshift-space remap  [a7]", 12, timed(plan, lambda: plan.rsd_remap(f2, f3, None, zgrid, 100.0, out), reps=3))
    add("apply_transfer_fn from delta_k  [a8]", 24,
        timed(plan, lambda: plan.spectrum_to_field(spec, f3, flags=F.F_FILTER)))
    add("binned P(k) of a field  [a10]", 20, timed(plan, lambda: plan.field_to_spectrum(field, want_pk=True)))
    add("binned P(k) + l=2,4 multipoles  [a10 ext]", 20,
        timed(plan, lambda: plan.field_to_spectrum(field, want_pk=True, poles=True)))
    add("binned P(k) from stored delta_k  [a10]", 4, timed(plan, lambda: plan.pk_from_spectrum(spec)))
    if N <= 1024:
        x = np.arange(N) - N / 2.
        s = 1.5 + 4.0 * np.arange(N) / N
        beam = plan.upload_f32(np.exp(-0.5 * (x[:, None, None] ** 2 + x[None, :, None] ** 2) / s[None, None, :] ** 2)
                               .astype(np.float32)) if N <= 512 else f3
        add("beam spectrum set-up (once per beam)  [a11]", 44, timed(plan, lambda: plan.beam_set(beam), reps=2))
        add("beam convolve_fft, cached beam spectrum  [a11]", 56,
            timed(plan, lambda: plan.beam_convolve(None, field, out), reps=3))
    u = plan.upload(np.random.default_rng(1).random(n3))
    counts = plan.alloc(n3 * 4)
    nbar = np.array([1e-3], np.float32)
    bias = np.array([1.0], np.float32)
    add("halo counts (Poisson inversion)  [a12]", 16,
        timed(plan, lambda: plan.halo_counts(field, nbar, 0, bias, 0, False, 0.0, u, counts), reps=3))
    # data-cube steps either side of the beam
    amps = np.random.default_rng(2).uniform(5., 50., (N, N)).astype(np.float32)
    alpha = np.random.default_rng(3).normal(-2.5, 0.3, (N, N)).astype(np.float32)
    l2f = np.log2(np.linspace(700., 1100., N) / 130.).astype(np.float32)
    add("foreground cube amps*(nu/nu0)^alpha  [f2]", 4, timed(plan, lambda: plan.fg_cube(amps, alpha, l2f, out), reps=3))
    add("foreground cube added to a cube  [f2]", 8, timed(plan, lambda: plan.fg_cube(amps, alpha, l2f, out, accumulate=True), reps=3))
    sig = np.linspace(0.5, 1.5, N)
    add("radiometer noise, Philox  [f2]", 4, timed(plan, lambda: plan.radiometer_noise(sig, out, None, seed=3), reps=3))
    add("radiometer noise added to a cube  [f2]", 8, timed(plan, lambda: plan.radiometer_noise(sig, out, None, seed=3, accumulate=True), reps=3))
    # PCA foreground filter, float64 device steps
    if N <= 1024:
        rng = np.random.default_rng(4)
        cube64 = plan.alloc(n3 * 8)
        clean64 = plan.alloc(n3 * 8)
        plan.lib.fb_convert_f32_to_f64(plan.h, _lib._ptr(field), _lib._ptr(cube64), n3)
        ms = timed(plan, lambda: plan.pca_covariance(cube64), reps=2)
        flops = 2.0 * (N * (N + 64) / 2.0) * N * N                 # upper-triangle tiles incl. the diagonal ones
        add("PCA covariance, float64, %.1f TFLOP/s FP64  [f3]" % (flops / (ms * 1e-3) / 1e12), 8, ms)
        mean, cov = plan.pca_covariance(cube64)
        w, v = np.linalg.eigh(cov)
        U = np.ascontiguousarray(v[:, ::-1][:, :4])
        add("PCA projection (4 modes), float64  [f3]", 16, timed(plan, lambda: plan.pca_project(cube64, mean, U, clean64), reps=2))
        del cube64, clean64
    # halo catalogue from the counts just drawn: 3 passes over the counts + 24 B per halo
    nh = plan.halo_catalogue(counts)
    cat = plan.alloc(max(nh, 1) * 24)
    ms = timed(plan, lambda: plan.halo_catalogue(counts, None, cat, nh), reps=3)
    add("halo catalogue, %.1f M halos  [f1]" % (nh / 1e6), 12 + 24.0 * nh / n3, ms)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(N=N, hbm_peak_GBs=peak, rows=rows), open(os.path.join(ROOT, "gpurun_out", "bench_all_%d.json" % N),
                                                            "w"), indent=1)


if __name__ == "__main__":
    main()
