"""
Parity at BASELINE.json's full single-GPU size (1024^3) through size-independent properties --
the CPU oracle cannot hold this size (the reference needs ~140 GB of float64 temporaries):
  * Parseval (box.py:944-946): sum(delta_x^2) N^3 == sum |delta_k|^2 (all modes, Hermitian weights);
  * the fused P(k) of realise == P(k) re-measured from the realised field by the forward transform
    (bin populations bit-identical, moments to 1e-5) and == P(k) from the stored spectrum;
  * linearity: realise with scale=2 is exactly 2x; filter of ones == unfiltered;
  * bin populations are independent of the data: equal to the 512^3-cross-checked closed form
    sum(count) == N^3, and to the counts of a second seed;
  * inverse(forward(field)) == field.
"""
import numpy as np
import pytest

from fastbox_b200 import _lib
from fastbox_b200 import kspace as ks

from _util import TOL, pk_function

pytestmark = pytest.mark.gpu
F = _lib


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))


import os


@pytest.mark.parametrize("N", [1024] + ([2048] if os.environ.get("FB_TEST_2048") else []))
def test_full_size_properties(gpu, N):
    L = 2000.0
    plan = _lib.Plan(N, L, L, L)
    _, pkf = pk_function(0.8)
    mode, tab, l0, dl = ks.choose_sqrt_pk_table(pkf, N, L, L, L, N ** 6. / L ** 3)
    plan.set_sqrt_pk(tab, mode, l0, dl)
    edges = ks.pk_bin_edges(2 * np.pi / L, 2 * np.pi * np.sqrt(3.) * N / L, 50)
    plan.set_pk_bins(ks.bin_thresholds(edges))
    n3 = N ** 3
    field = plan.alloc(n3 * 4)
    spec = plan.alloc((N // 2 + 1) * N * N * 8)
    res, sums = plan.realise(None, None, seed=2024, flags=F.F_SQRTPK, field_out=field, spec_out=spec, want_pk=True)
    # every mode is counted exactly once (Hermitian multiplicities) -- integer, bit exact
    assert int(res["count"].sum()) == n3
    # Parseval
    total_k = float(res["sum1"].sum()) * (N ** 6. / L ** 3)
    assert abs(sums[1] * n3 / total_k - 1) < 1e-5
    assert abs(sums[0]) / np.sqrt(sums[1] * n3) < 1e-4                 # zero mean (P(0) = 0, box.py:167)
    # forward transform of the realised field gives the same binned spectrum
    fwd = plan.field_to_spectrum(field, want_pk=True)
    assert np.array_equal(fwd["count"], res["count"])
    m = res["count"][:50] > 0
    floor = 1e-12 * np.abs(res["sum1"]).max()          # bin 0 holds only the DC mode, whose power is exactly 0
    assert np.all(np.abs(fwd["sum1"][:50][m] - res["sum1"][:50][m]) <= 2 * TOL * np.abs(res["sum1"][:50][m]) + floor)
    st = plan.pk_from_spectrum(spec)
    assert np.array_equal(st["count"], res["count"])
    assert np.all(np.abs(st["sum1"][:50][m] - res["sum1"][:50][m]) <= 1e-6 * np.abs(res["sum1"][:50][m]) + floor)
    # second seed: identical bin populations (they depend on the grid only)
    f2 = plan.alloc(n3 * 4)
    res2, _ = plan.realise(None, None, seed=7, flags=F.F_SQRTPK, field_out=f2, want_pk=True)
    assert np.array_equal(res2["count"], res["count"])
    # measured P(k) follows the input spectrum: mean power per bin ~ P(k_c) within sample variance
    kc, pk, err = ks.moments_to_spectrum(edges, res["count"], res["sum1"], res["sum2"])
    good = res["count"][1:50] > 1000
    assert np.all(np.abs(pk[good] / pkf(kc[good]) - 1) < 0.25)         # bin-centre vs bin-average, coarse
    # linearity in `scale` and round trip inverse(forward(x)) == x on a slab (host copies are 4 GB each)
    plan.spectrum_to_field(spec, f2, scale=2.0)
    h1 = plan.download(field, (8, N, N), np.float32)
    h2 = plan.download(f2, (8, N, N), np.float32)
    assert _rel(h2, 2.0 * h1) < 1e-6
    plan.field_to_spectrum(field, spec_out=spec)
    plan.spectrum_to_field(spec, f2)
    h2 = plan.download(f2, (8, N, N), np.float32)
    assert _rel(h2, h1) < TOL
    plan.close()
