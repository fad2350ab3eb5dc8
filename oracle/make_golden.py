"""
TEST INFRASTRUCTURE ONLY -- regenerate ``tests/golden/*.npz`` by running the
UNMODIFIED reference (``/root/reference/fastbox``) through ``ref_loader``.

    python oracle/make_golden.py

The fixtures hold the reference's outputs for the canonical inputs of its own
tests (seeds 11/14, boxes 16^3 @ 100 Mpc, cuboid (1e2,2e2,1e3), 32^3 @ 1 Gpc;
filter of tests/test_box.py:88-90; sigma_nl RSD; Gaussian beam cube).  Inputs
are re-derived in the tests from the stored seed with NumPy's legacy global
generator, exactly as the reference draws them (box.py:174-175, 418).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

CASES = [
    dict(name="n16_cubic", N=16, scale=(1e2, 1e2, 1e2), seed=11, redshift=0.0),
    dict(name="n16_cuboid", N=16, scale=(1e2, 2e2, 1e3), seed=11, redshift=1.0),
    dict(name="n32_gpc", N=32, scale=1e3, seed=14, redshift=0.8),
]


def transfer_fn(k_perp, k_par):                      # tests/test_box.py:88-90
    return (1. - np.exp(-0.5 * (k_par / 0.001) ** 2.)) * np.exp(-0.5 * (k_perp / 0.1) ** 2.)


def beam_cube(N):
    x = np.arange(N) - N / 2.
    s = 1.5 + 0.05 * np.arange(N)
    return np.exp(-0.5 * (x[:, None, None] ** 2 + x[None, :, None] ** 2) / s[None, None, :] ** 2)


def run_case(c):
    box_m = ref_loader.load("box")
    beams_m = ref_loader.load("beams")
    halos_m = ref_loader.load("halos")
    tracers_m = ref_loader.load("tracers")
    N = c["N"]
    np.random.seed(c["seed"])
    box = box_m.CosmoBox(cosmo=box_m.default_cosmo, box_scale=c["scale"], nsamp=N,
                         redshift=c["redshift"], realise_now=False)
    box.realise_density()
    out = dict(N=N, seed=c["seed"], redshift=c["redshift"],
               scale=np.atleast_1d(np.array(c["scale"], dtype=np.float64)),
               L=np.array([box.Lx, box.Ly, box.Lz]), boxfactor=box.boxfactor,
               kmin=box.kmin, kmax=box.kmax,
               delta_x=box.delta_x, delta_k_half=box.delta_k[:N // 2 + 1])
    for nb in (20, 50):
        kc, pk, err = box.binned_power_spectrum(nbins=nb)
        out["pk%d_k" % nb], out["pk%d_p" % nb], out["pk%d_e" % nb] = kc, pk, err
        bins = np.logspace(np.log10(box.kmin), np.log10(box.kmax), nb)
        idx = np.digitize(box.k.flatten(), bins)
        out["pk%d_counts" % nb] = np.bincount(idx, minlength=nb + 1)
    s1, s2 = box.test_parseval()
    out["parseval"] = np.array([s1, s2])
    vk = box.realise_velocity()
    vz = np.fft.ifftn(vk[2]).real
    out["vel_z"] = vz
    out["vel_x"] = np.fft.ifftn(vk[0]).real
    out["transfer"] = box.apply_transfer_fn(box.delta_k, transfer_fn).real
    out["transfer_imag_max"] = np.abs(box.apply_transfer_fn(box.delta_k, transfer_fn).imag).max()
    out["smooth8"] = box.smooth_field(box.delta_k, 8.0).real
    tr = tracers_m.HITracer(box)
    bias = tr.bias_HI()
    out["bias_HI"] = bias
    out["Tb"] = tr.signal_amplitude()
    out["lognormal"] = box.lognormal(box.delta_x * bias)
    out["rsd0"] = box.redshift_space_density(delta_x=out["lognormal"], velocity_z=vz, sigma_nl=0.)
    np.random.seed(c["seed"] + 100)
    out["rsd120"] = box.redshift_space_density(delta_x=out["lognormal"], velocity_z=vz, sigma_nl=120.)

    class GaussBeam(beams_m.BeamModel):
        def beam_cube(self, pol=None):
            return beam_cube(N)
    out["beam_conv"] = GaussBeam(box).convolve_fft(out["rsd0"])
    # mean halo count (everything before the Poisson draw, halos.py:91-113)
    hd = halos_m.HaloDistribution(box, (1e12, 1e15), 10)
    real_poisson = np.random.poisson
    np.random.poisson = lambda lam: lam
    try:
        out["halo_mean_ln"] = hd.halo_count_field(box.delta_x, nbar=1e-3, bias=1.2, lognormal=True)
        out["halo_mean_lin"] = hd.halo_count_field(box.delta_x, nbar=np.linspace(1e-3, 2e-3, N),
                                                   bias=1.2, lognormal=False)
    finally:
        np.random.poisson = real_poisson
    out["freq"] = box.freq_array()
    out["pix_x"] = box.pixel_array()[0]
    out["z_grid"] = box.z
    import pyccl as ccl
    out["Hz"] = 100. * box.cosmo['h'] * ccl.h_over_h0(box.cosmo, box.scale_factor)
    return out


def run_catalogue():
    """HaloDistribution.realise_halo_catalogue of the unmodified reference (halos.py:120-176)."""
    box_m = ref_loader.load("box")
    halos_m = ref_loader.load("halos")
    out = {}
    for name, N, scale, lam in (("sparse", 16, (1.0, 0.5, 0.25), 0.05), ("dense", 8, 0.1, 3.0), ("mixed", 32, 0.3, 0.4)):
        box = box_m.CosmoBox(cosmo=box_m.default_cosmo, box_scale=scale, nsamp=N, redshift=0.4, realise_now=False)
        hd = halos_m.HaloDistribution(box, (1e12, 1e15), 10)
        np.random.seed(4242 + N)
        counts = np.random.poisson(lam=lam, size=(N, N, N))
        if name == "mixed":
            counts[3, 5, 7] = 40                       # one rich voxel: many distinct count values
            counts[31, 31, 31] = 17
        out[name + "_counts"] = counts.astype(np.int32)
        out[name + "_L"] = np.array([box.Lx, box.Ly, box.Lz])
        out[name + "_cat"] = hd.realise_halo_catalogue(counts, scatter=False)
        np.random.seed(99 + N)
        out[name + "_cat_scatter"] = hd.realise_halo_catalogue(counts, scatter=True)
        out[name + "_scatter_seed"] = 99 + N
    return out


def run_cube():
    """ForegroundModel (foregrounds.py:48-174) and NoiseModel (noise.py:25-75) of the unmodified reference."""
    box_m = ref_loader.load("box")
    fg_m = ref_loader.load("foregrounds")
    noise_m = ref_loader.load("noise")
    N = 16
    box = box_m.CosmoBox(cosmo=box_m.default_cosmo, box_scale=(4e2, 3e2, 2e2), nsamp=N, redshift=0.8,
                         realise_now=False)
    fg = fg_m.ForegroundModel(box)
    out = dict(N=N, scale=np.array([4e2, 3e2, 2e2]), redshift=0.8)
    np.random.seed(77)
    out["amps"] = fg.realise_foreground_amp(amp=57., beta=-1.1, monopole=10., smoothing_scale=4.)   # example_endtoend.py:60-62
    out["amps_nosmooth"] = fg.realise_foreground_amp(amp=57., beta=-1.1, monopole=10.)
    out["alpha"] = fg.realise_spectral_index(mean_spec_idx=-2.07, std_spec_idx=0.2, smoothing_scale=15.)
    out["fg_cube_map"] = fg.construct_cube(out["amps"], out["alpha"], freq_ref=130.)
    out["fg_cube_scalar"] = fg.construct_cube(out["amps"], -2.7, freq_ref=130.)
    np.random.seed(78)
    out["noise"] = noise_m.NoiseModel(box).realise_radiometer_noise(Tinst=18., tp=2.5, fov=1., Ndish=64)
    out["freqs"] = box.freq_array()
    out["ang_x"] = box.pixel_array()[0]
    filters_m = ref_loader.load_pkg("filters")
    out["data_cube"] = out["fg_cube_map"] + out["noise"]
    out["mean_filtered"] = filters_m.mean_spectrum_filter(out["data_cube"])      # filters.py:35-55
    # PCA cleaning (filters.py:93-183) of foregrounds + noise + a small "signal" (example_endtoend.py:101-105)
    np.random.seed(79)
    sig = 0.1 * np.random.normal(0., 1., out["data_cube"].shape)
    out["pca_cube"] = out["data_cube"] + sig
    for nm in (2, 4):
        out["pca_clean%d" % nm] = np.real(filters_m.pca_filter(out["pca_cube"], nmodes=nm))
    c, U, a = filters_m.pca_filter(out["pca_cube"], nmodes=3, return_filter=True)
    out["pca_clean3"], out["pca_U3"], out["pca_amps3"] = np.real(c), np.real(U), np.real(a)
    out["pca_clean3_pl"] = np.real(filters_m.pca_filter(out["pca_cube"], nmodes=3, fit_powerlaw=True))
    # foreground-dominated cube: foregrounds ~3e3 x the signal
    out["pca_cube_fg"] = 1e3 * out["fg_cube_map"] + out["noise"] + sig
    out["pca_fg_clean3"] = np.real(filters_m.pca_filter(out["pca_cube_fg"], nmodes=3))
    return out


def run_rsd_methods():
    """redshift_space_density(method='nearest') of the unmodified reference (box.py:403-405, 433-437) on the
    inputs already held by the per-case fixtures (their log-normal field and v_z), without and with sigma_nl."""
    box_m = ref_loader.load("box")
    out = {}
    for c in CASES:
        g = np.load(os.path.join(ROOT, "tests", "golden", c["name"] + ".npz"))
        box = box_m.CosmoBox(cosmo=box_m.default_cosmo, box_scale=c["scale"], nsamp=c["N"],
                             redshift=c["redshift"], realise_now=False)
        out[c["name"] + "_nearest0"] = box.redshift_space_density(delta_x=g["lognormal"], velocity_z=g["vel_z"],
                                                                  sigma_nl=0., method="nearest")
        np.random.seed(c["seed"] + 100)
        out[c["name"] + "_nearest120"] = box.redshift_space_density(delta_x=g["lognormal"], velocity_z=g["vel_z"],
                                                                    sigma_nl=120., method="nearest")
    return out


def main():
    warnings.simplefilter("ignore")
    dest = os.path.join(ROOT, "tests", "golden")
    os.makedirs(dest, exist_ok=True)
    if "--rsd-methods" in sys.argv:                     # needs the per-case fixtures: with --all it runs last
        path = os.path.join(dest, "rsd_methods.npz")
        np.savez_compressed(path, **run_rsd_methods())
        print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024.))
        return
    if "--cube" in sys.argv or "--all" in sys.argv:
        path = os.path.join(dest, "fg_noise_cube.npz")
        np.savez_compressed(path, **run_cube())
        print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024.))
        if "--all" not in sys.argv and "--catalogue" not in sys.argv:
            return
    if "--catalogue" in sys.argv or "--all" in sys.argv:
        path = os.path.join(dest, "halo_catalogue.npz")
        np.savez_compressed(path, **run_catalogue())
        print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024.))
        if "--all" not in sys.argv:
            return
    for c in CASES:
        out = run_case(c)
        path = os.path.join(dest, c["name"] + ".npz")
        np.savez_compressed(path, **out)
        print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024.))
    if "--all" in sys.argv:
        path = os.path.join(dest, "rsd_methods.npz")
        np.savez_compressed(path, **run_rsd_methods())
        print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024.))


if __name__ == "__main__":
    main()
