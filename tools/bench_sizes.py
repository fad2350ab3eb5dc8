"""
Headline pipeline (realise + k_perp/k_par filter + binned P(k)) at several grid sizes on one GPU.
HBM-resident noise (28 B/cell) up to 1024^3; Philox noise (20 B/cell) up to 2048^3.

    python tools/bench_sizes.py
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


def main():
    peak = 6541.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rows = []
    for N in (128, 256, 512, 1024, 2048):
        L = 2000.0 * N / 1024
        plan = _lib.Plan(N, L, L, L)
        _, pkf = pk_function(0.8)
        with np.errstate(all="ignore"):
            mode, tab, l0, dl = ks.choose_sqrt_pk_table(pkf, N, L, L, L, N ** 6. / L ** 3)
        plan.set_sqrt_pk(tab, mode, l0, dl)
        ft = ks.filter_tables(transfer_fn, N, L, L, L)
        plan.set_filter(ft.tperp, ft.tpar, ft.tdense)
        plan.set_pk_bins(ks.bin_thresholds(ks.pk_bin_edges(2 * np.pi / L, 2 * np.pi * np.sqrt(3.) * N / L, 50)))
        n3 = N ** 3
        field = plan.alloc(n3 * 4)
        reps = 20 if N <= 512 else 5

        def run(fn):
            fn()
            plan.sync()
            plan.timer_start()
            for _ in range(reps):
                fn()
            return plan.timer_stop() / reps
        flags = F.F_SQRTPK | F.F_FILTER
        ms_p = run(lambda: plan.realise(None, None, seed=1, flags=flags, field_out=field, want_pk=True))
        row = dict(N=N, sqrtp_mode=int(mode), philox_ms=ms_p, philox_Gcells_s=n3 / ms_p / 1e6,
                   philox_frac_hbm=20.0 * n3 / (ms_p * 1e-3) / 1e9 / peak)
        if N <= 1024:
            re = plan.alloc(n3 * 4)
            im = plan.alloc(n3 * 4)
            plan.affine(re, n3, 0.0, 0.5)
            plan.affine(im, n3, 0.0, -0.25)
            ms_n = run(lambda: plan.realise(re, im, flags=flags, field_out=field, want_pk=True))
            row.update(noise_ms=ms_n, noise_Gcells_s=n3 / ms_n / 1e6, noise_frac_hbm=28.0 * n3 / (ms_n * 1e-3) / 1e9 / peak)
        rows.append(row)
        print(json.dumps(row), flush=True)
        plan.close()
        del field
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(hbm_peak_GBs=peak, rows=rows), open(os.path.join(ROOT, "gpurun_out", "bench_sizes.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
