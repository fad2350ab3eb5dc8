"""
GPU parity for the data-cube steps either side of the beam convolution:
ForegroundModel.construct_cube (foregrounds.py:152-174) and NoiseModel.realise_radiometer_noise
(noise.py:25-75), against the golden vectors of the unmodified reference and the oracle.  Floating
point: 1e-5 relative L2 in float32 (north_star tolerance), written below as TOL.
"""
import numpy as np
import pytest

import fastbox_b200 as fb
from fastbox_b200 import _lib
from fastbox_b200.box import CosmoBox, default_cosmo
from oracle import restate as R
from _util import TOL, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _box(g):
    return CosmoBox(cosmo=default_cosmo, box_scale=tuple(g["scale"]), nsamp=int(g["N"]), redshift=float(g["redshift"]),
                    realise_now=False)


def test_foreground_model_matches_reference_golden(gpu):
    g = load_golden("fg_noise_cube")
    box = _box(g)
    fg = fb.foregrounds.ForegroundModel(box)
    np.random.seed(77)                                   # same draws as oracle/make_golden.py:run_cube
    amps = fg.realise_foreground_amp(amp=57., beta=-1.1, monopole=10., smoothing_scale=4.)
    amps2 = fg.realise_foreground_amp(amp=57., beta=-1.1, monopole=10.)
    alpha = fg.realise_spectral_index(mean_spec_idx=-2.07, std_spec_idx=0.2, smoothing_scale=15.)
    assert np.array_equal(amps, g["amps"]) and np.array_equal(amps2, g["amps_nosmooth"])
    assert np.array_equal(alpha, g["alpha"])
    cube = fg.construct_cube(amps, alpha, freq_ref=130.)
    assert cube.dtype == np.float64 and cube.shape == g["fg_cube_map"].shape
    assert rel_l2(cube, g["fg_cube_map"]) < TOL
    assert np.max(np.abs(cube / g["fg_cube_map"] - 1)) < 1e-5            # every voxel, not only in the mean
    assert rel_l2(fg.construct_cube(amps, -2.7, freq_ref=130.), g["fg_cube_scalar"]) < TOL
    # extension: foregrounds added to an existing cube in the same pass
    base = np.random.default_rng(1).standard_normal(cube.shape)
    assert rel_l2(fg.construct_cube(amps, alpha, add_to=base), base + g["fg_cube_map"]) < TOL
    with pytest.raises(ValueError):
        fg.construct_cube(amps[:4], alpha)


def test_noise_model_matches_reference_golden(gpu):
    g = load_golden("fg_noise_cube")
    box = _box(g)
    nm = fb.noise.NoiseModel(box)
    assert np.allclose(nm.radiometer_rms(18., 2.5, 1., 64),
                       R.radiometer_rms_port(g["freqs"], g["ang_x"], 18., 2.5, 1., 64), rtol=1e-12)
    np.random.seed(78)
    noise = nm.realise_radiometer_noise(Tinst=18., tp=2.5, fov=1., Ndish=64)
    assert noise.dtype == np.float64 and rel_l2(noise, g["noise"]) < TOL
    # device-drawn normals: reproducible, equal to the restated Philox stream, right rms per channel
    a = nm.realise_radiometer_noise(18., 2.5, 1., 64, seed=5)
    b = nm.realise_radiometer_noise(18., 2.5, 1., 64, seed=5)
    assert np.array_equal(a, b)
    sig = nm.radiometer_rms(18., 2.5, 1., 64)
    ref = R.radiometer_noise_port(sig, R.philox_noise_cube(5, int(g["N"])))
    assert rel_l2(a, ref) < TOL and np.max(np.abs(a - ref)) < 1e-3 * sig.max()
    c = nm.realise_radiometer_noise(18., 2.5, 1., 64, seed=6)
    assert not np.array_equal(a, c)


@pytest.mark.parametrize("N", [8, 64, 256])
def test_cube_kernels_vs_oracle(gpu, N):
    rng = np.random.default_rng(N)
    plan = _lib.Plan(N, 1e3, 1e3, 1e3)
    amps = rng.uniform(5., 50., (N, N))
    idx = rng.normal(-2.5, 0.3, (N, N))
    freqs = np.linspace(700., 1100., N)
    out = np.empty((N, N, N), np.float32)
    plan.fg_cube(amps, idx, np.log2(freqs / 130.), out)
    ref = R.fg_construct_cube_port(amps.astype(np.float32), idx.astype(np.float32), freqs)
    assert rel_l2(out, ref) < TOL and np.max(np.abs(out / ref - 1)) < 1e-5
    sig = rng.uniform(0.1, 2.0, N)
    normals = rng.standard_normal((N, N, N)).astype(np.float32)
    plan.radiometer_noise(sig, out, normals)
    assert rel_l2(out, R.radiometer_noise_port(sig, normals)) < TOL
    acc = out.copy()
    plan.radiometer_noise(sig, acc, normals, accumulate=True)           # host cube: upload, add, download
    assert rel_l2(acc, 2 * R.radiometer_noise_port(sig, normals)) < TOL
    plan.radiometer_noise(sig, out, None, seed=N)
    ref = R.radiometer_noise_port(sig, R.philox_noise_cube(N, N))
    # fast-math Box-Muller: typical absolute error 5e-7 of a unit normal; the rare u1 -> 1 draws (|n| ~ 1e-3)
    # carry the absolute error of __logf near 1, hence the looser bound on the maximum
    assert rel_l2(out, ref) < TOL and np.max(np.abs(out - ref)) < 1e-3 * sig.max()
    plan.close()


def test_mean_spectrum_filter(gpu):
    """filters.mean_spectrum_filter (filters.py:35-55): golden of the unmodified reference, then a cube with a
    large monopole (the per-channel sums are float64, so the residual keeps its digits)."""
    g = load_golden("fg_noise_cube")
    out, mean = fb.filters.mean_spectrum_filter(g["data_cube"], return_mean=True)
    assert out.dtype == np.float64 and out.shape == g["mean_filtered"].shape
    ref_mean = g["data_cube"].reshape(-1, out.shape[-1]).mean(axis=0)
    assert np.allclose(mean, ref_mean, rtol=1e-6, atol=0)
    # the cube is float32 on the device: tolerance relative to the cube, not to the (much smaller) residual
    scale = np.sqrt(np.mean(g["data_cube"] ** 2))
    assert np.sqrt(np.mean((out - g["mean_filtered"]) ** 2)) < TOL * scale
    rng = np.random.default_rng(3)
    N = 64
    cube = (1e4 * np.linspace(1.0, 2.0, N)[None, None, :] + rng.standard_normal((N, N, N))).astype(np.float32)
    out = fb.filters.mean_spectrum_filter(cube)
    ref = R.mean_spectrum_filter_port(cube.astype(np.float64))
    assert np.max(np.abs(out - ref)) < 1e-6                # float64 subtraction: only the final float32 rounding of O(1) values
    assert abs(out.mean()) < 1e-7
    with pytest.raises(ValueError):
        fb.filters.mean_spectrum_filter(np.zeros((4, 4, 8)))


def test_pca_filter_matches_reference_golden(gpu):
    """filters.pca_filter (filters.py:93-183), float64 on the device, against the unmodified reference's output.
    The cleaned cube is invariant to the sign of the eigenvectors; U_fg / fg_amps are compared up to that sign."""
    g = load_golden("fg_noise_cube")
    for nm in (2, 4):
        out = fb.filters.pca_filter(g["pca_cube"], nmodes=nm)
        assert out.dtype == np.float64 and rel_l2(out, g["pca_clean%d" % nm]) < TOL
    c, U, a = fb.filters.pca_filter(g["pca_cube"], nmodes=3, return_filter=True)
    assert rel_l2(c, g["pca_clean3"]) < 1e-10                     # float64 path: far inside the 1e-5 bar
    assert U.shape == g["pca_U3"].shape and a.shape == g["pca_amps3"].shape
    sign = np.sign(np.sum(U * g["pca_U3"], axis=0))
    assert np.allclose(U * sign, g["pca_U3"], atol=1e-9) and np.allclose(a * sign[:, None], g["pca_amps3"], atol=1e-8)
    assert rel_l2(fb.filters.pca_filter(g["pca_cube"], nmodes=3, fit_powerlaw=True), g["pca_clean3_pl"]) < TOL
    # foreground-dominated cube (rms 246 against a cleaned rms of 0.09): float32 storage alone would be off by 1e-4
    out = fb.filters.pca_filter(g["pca_cube_fg"], nmodes=3)
    assert rel_l2(out, g["pca_fg_clean3"]) < 1e-9
    with pytest.raises(ValueError):
        fb.filters.pca_filter(g["pca_cube"], nmodes=0)


@pytest.mark.parametrize("N,tile,kt", [(8, 0, 0), (64, 0, 0), (128, 0, 0), (128, 128, 8), (256, 0, 0), (256, 128, 8),
                                       (256, 64, 0), (512, 0, 0)])
def test_pca_covariance_and_projection_vs_numpy(gpu, monkeypatch, N, tile, kt):
    """tile / kt select the covariance kernel (FB_PCA_TILE: 64 x 64 SIMT tiles or 128 x 128 tiles on the FP64 tensor
    path; FB_PCA_KT: panel depth); 0 = the default for that size (128 x 128, 16-pixel panels from 256 channels on)."""
    if tile:
        monkeypatch.setenv("FB_PCA_TILE", str(tile))
    if kt:
        monkeypatch.setenv("FB_PCA_KT", str(kt))
    rng = np.random.default_rng(N)
    nu = np.linspace(1.0, 2.0, N)
    cube = (50.0 * rng.uniform(0.5, 1.5, (N, N, 1)) * nu[None, None, :] ** -2.7
            + 5.0 * rng.standard_normal((N, N, 1)) * nu[None, None, :] ** -1.0
            + 0.05 * rng.standard_normal((N, N, N)))
    plan = _lib.Plan(N, 1., 1., 1.)
    d = plan.upload(cube)
    mean, cov = plan.pca_covariance(d)
    x = cube.reshape(-1, N)
    assert np.allclose(mean, x.mean(axis=0), rtol=1e-13)
    ref_cov = np.cov(x.T)
    assert np.max(np.abs(cov - ref_cov)) < 1e-12 * np.max(np.abs(ref_cov)) and np.array_equal(cov, cov.T)
    assert rel_l2(fb.filters.pca_filter(cube, nmodes=2), R.pca_filter_port(cube, 2)) < 1e-8
    plan.close()
