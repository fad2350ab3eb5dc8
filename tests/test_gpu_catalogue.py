"""
GPU parity: fb_halo_catalogue (counts per voxel -> catalogue of positions, halos.py:120-176) against
the golden vectors of the unmodified reference and against the oracle; integer / index work, so the
bar is bit-exact.  Through the C ABI (Plan) and through the drop-in class.
"""
import numpy as np
import pytest

import fastbox_b200 as fb
from fastbox_b200 import _lib
from fastbox_b200.box import CosmoBox, default_cosmo
from oracle import restate as R
from _util import load_golden

pytestmark = pytest.mark.gpu


def _catalogue(plan, counts, uniforms=None):
    c = np.ascontiguousarray(counts, dtype=np.int32)
    nh = plan.halo_catalogue(c)
    cat = np.full((nh, 3), np.nan)
    if nh:
        assert plan.halo_catalogue(c, uniforms, cat, nh) == nh
    return cat


@pytest.mark.parametrize("name", ["sparse", "dense", "mixed"])
def test_catalogue_matches_reference_golden(gpu, name):
    g = load_golden("halo_catalogue")
    counts, L = g[name + "_counts"], g[name + "_L"]
    plan = _lib.Plan(counts.shape[0], *L)
    cat = _catalogue(plan, counts)
    assert cat.shape == g[name + "_cat"].shape and np.array_equal(cat, g[name + "_cat"])
    np.random.seed(int(g[name + "_scatter_seed"]))
    u = np.random.uniform(0., 1. - 1e-8, cat.size).reshape(cat.shape)
    assert np.array_equal(_catalogue(plan, counts, u), g[name + "_cat_scatter"])


@pytest.mark.parametrize("N,lam", [(8, 0.0), (8, 0.01), (16, 2.5), (64, 0.02), (64, 1.0), (128, 0.3)])
def test_catalogue_matches_oracle(gpu, N, lam):
    rng = np.random.default_rng(N + int(100 * lam))
    counts = rng.poisson(lam, (N, N, N)).astype(np.int32)
    if N == 64:
        counts[0, 0, 0], counts[N - 1, N - 1, N - 1], counts[5, 6, 7] = 1023, 300, 299     # many keys, both ends of the grid
    L = (300.0, 200.0, 123.4)
    plan = _lib.Plan(N, *L)
    cat = _catalogue(plan, counts)
    ref = R.halo_catalogue_port(counts, *L)
    assert cat.shape == ref.shape and np.array_equal(cat, ref)
    if cat.size:
        u = rng.random(cat.shape)
        assert np.array_equal(_catalogue(plan, counts, u), R.halo_catalogue_port(counts, *L, uniforms=u))


def test_catalogue_error_paths(gpu):
    plan = _lib.Plan(8, 1., 1., 1.)
    c = np.zeros((8, 8, 8), np.int32)
    c[1, 2, 3] = 1024                                       # above the supported per-voxel maximum
    with pytest.raises(RuntimeError, match="exceeds"):
        plan.halo_catalogue(c)
    c[1, 2, 3] = -1
    with pytest.raises(RuntimeError, match="negative"):
        plan.halo_catalogue(c)
    c[1, 2, 3] = 3
    with pytest.raises(RuntimeError, match="buffer holds"):
        plan.halo_catalogue(c, None, np.empty((2, 3)), 2)


def test_drop_in_class_reproduces_reference_stream(gpu):
    """Same np.random state -> same catalogue as the reference class (offsets drawn as halos.py:166)."""
    g = load_golden("halo_catalogue")
    counts = g["mixed_counts"]
    box = CosmoBox(cosmo=default_cosmo, box_scale=0.3, nsamp=32, redshift=0.4, realise_now=False)
    assert np.array_equal([box.Lx, box.Ly, box.Lz], g["mixed_L"])
    hd = fb.halos.HaloDistribution(box, (1e12, 1e15), 10)
    assert np.array_equal(hd.realise_halo_catalogue(counts.astype(np.int64)), g["mixed_cat"])
    np.random.seed(int(g["mixed_scatter_seed"]))
    assert np.array_equal(hd.realise_halo_catalogue(counts, scatter=True), g["mixed_cat_scatter"])
    with pytest.raises(ValueError):
        hd.realise_halo_catalogue(counts, scatter=True, scatter_type="gauss")
