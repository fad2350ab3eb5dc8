"""GPU probe: strided-copy bandwidth vs chunk width, and per-pass timings of the realise pipeline."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastbox_b200 import _lib  # noqa: E402
from fastbox_b200 import kspace as ks  # noqa: E402


def main():
    out = {}
    print(_lib.device_info(0))
    plan = _lib.Plan(64, 1e3, 1e3, 1e3)
    for chunk in (32, 64, 128, 256, 512, 8192):
        g = plan.bench_strided_copy(4 << 30, chunk, 5)
        out["copy_chunk_%d" % chunk] = g
        print("strided copy chunk %5d B: %8.1f GB/s" % (chunk, g), flush=True)
    plan.close()
    sizes = [int(x) for x in (sys.argv[1:] or ["256", "512", "1024"])]
    for N in sizes:
        L = 2000.0
        plan = _lib.Plan(N, L, L, L)
        n2 = np.arange(3 * (N // 2) ** 2 + 1, dtype=np.float64)
        k = 2 * np.pi * np.sqrt(n2) / L
        with np.errstate(all="ignore"):
            lut = np.where(k > 0, 1e4 * (k / 0.02) / (1 + (k / 0.02) ** 2.5), 0.0)
        bf = N ** 6. / L ** 3
        plan.set_sqrt_pk(np.sqrt(lut * bf).astype(np.float32), 1)
        kmin, kmax = 2 * np.pi / L, 2 * np.pi * np.sqrt(3.) * N / L
        plan.set_pk_bins(ks.bin_thresholds(ks.pk_bin_edges(kmin, kmax, 50)))
        m = ks.mode_numbers(N).astype(np.float64)
        h = N // 2 + 1
        kperp = 2 * np.pi * np.sqrt((m[:h, None] / L) ** 2 + (m[None, :] / L) ** 2)
        kpar = 2 * np.pi * m / L
        plan.set_filter(np.exp(-0.5 * (kperp / 0.1) ** 2), 1. - np.exp(-0.5 * (kpar / 0.001) ** 2), None)
        field = plan.alloc(N ** 3 * 4)
        for label, flags, pk in (("philox", _lib.F_SQRTPK, False),
                                 ("philox+filter+pk", _lib.F_SQRTPK | _lib.F_FILTER, True)):
            best = None
            for it in range(4):
                plan.realise(None, None, seed=it, flags=flags, field_out=field, want_pk=pk)
                t = plan.last_timings(3)
                if best is None or sum(t) < sum(best):
                    best = t
            tot = sum(best)
            print("N=%d %-18s rows %.3f cols %.3f x %.3f total %.3f ms  -> %.1f Mcells/s, %.1f GB/s (20 B/cell)"
                  % (N, label, best[0], best[1], best[2], tot, N ** 3 / tot / 1e3, 20 * N ** 3 / tot / 1e6), flush=True)
            out["N%d_%s" % (N, label)] = best
        # noise from device-resident re/im
        if N <= 1024:
            rng = np.random.default_rng(0)
            re = plan.alloc(N ** 3 * 4)
            im = plan.alloc(N ** 3 * 4)
            # fill with something cheap: reuse the field buffer contents
            _lib.check(_lib.load().fb_copy(plan.h, re.ptr, field.ptr, N ** 3 * 4))
            _lib.check(_lib.load().fb_copy(plan.h, im.ptr, field.ptr, N ** 3 * 4))
            best = None
            for it in range(4):
                plan.realise(re, im, flags=_lib.F_SQRTPK | _lib.F_FILTER, field_out=field, want_pk=True)
                t = plan.last_timings(3)
                if best is None or sum(t) < sum(best):
                    best = t
            tot = sum(best)
            print("N=%d %-18s rows %.3f cols %.3f x %.3f total %.3f ms  -> %.1f Mcells/s, %.1f GB/s (28 B/cell)"
                  % (N, "noise+filter+pk", best[0], best[1], best[2], tot, N ** 3 / tot / 1e3,
                     28 * N ** 3 / tot / 1e6), flush=True)
            out["N%d_noise" % N] = best
            # forward P(k)
            best = None
            for it in range(4):
                plan.field_to_spectrum(field, want_pk=True)
                t = plan.last_timings(3)
                if best is None or sum(t) < sum(best):
                    best = t
            tot = sum(best)
            print("N=%d %-18s x %.3f cols %.3f rows %.3f total %.3f ms  -> %.1f Mcells/s, %.1f GB/s (20 B/cell)"
                  % (N, "forward P(k)", best[0], best[1], best[2], tot, N ** 3 / tot / 1e3,
                     20 * N ** 3 / tot / 1e6), flush=True)
            out["N%d_forward" % N] = best
            re.free(); im.free()
        field.free()
        plan.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
