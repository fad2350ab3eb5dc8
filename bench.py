#!/usr/bin/env python
"""
bench.py -- headline benchmark of the FastBox field-generation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S]

Metric (BASELINE.json): Mcells/s of the fused realise + anisotropic k_perp/k_par filter + binned
P(k) pipeline.  One "step" = one full pass over one synthetic N^3 box.

* N = 1 : 1024^3 (BASELINE.json configs[2] without the beam stage), white noise (re, im) float32
  RESIDENT IN HBM when the timed region starts (28 algorithmic B/cell).
  `value`  = device-resident throughput, CUDA events on the library stream.
  `e2e`    = the same call through the C ABI with HOST (pinned) buffers: noise H2D and
             field + P(k) D2H inside the timed region.
* N > 1 : 2048^3 (configs[4]) slab-decomposed over the ranks, counter-based Philox noise,
  one NCCL all-to-all per transform; max-over-ranks device time.
* --impl reference : the reference's own algorithm (NumPy port of fastbox/box.py, the oracle
  in oracle/restate.py: full complex128 numpy.fft c2c, masked per-bin loops) on the host cores,
  on a bounded box (the reference cannot hold 1024^3: ~140 GB of float64 temporaries).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Mcells/s realise+filter+P(k)"
DIST_XMODE_DEFAULT = 2   # exchange mechanism of the sharded run (fb_dist.cu); FB_DIST_XMODE overrides
BYTES_PER_CELL_HBM_NOISE = 28.0      # 8+4, 4+4, 4+4  
BYTES_PER_CELL_PHILOX = 20.0         # 0+4, 4+4, 4+4
PASS1_BYTES_PER_CELL = {"hbm": 12.0, "philox": 4.0}
Z_BOX = 0.8
NBINS = 50


def transfer_fn(k_perp, k_par):      # reference tests/test_box.py:88-90
    return (1. - np.exp(-0.5 * (k_par / 0.001) ** 2.)) * np.exp(-0.5 * (k_perp / 0.1) ** 2.)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.stop = threading.Event()
        self.th = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                self.samples.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for n, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def configure_plan(plan, N, L):
    """sqrt(P) LUT, filter tables, bins -- what CosmoBox would set up for this box."""
    from fastbox_b200 import kspace as ks
    from _util import pk_function
    _, pkf = pk_function(Z_BOX)
    bf = N ** 6. / L ** 3
    with np.errstate(all="ignore"):
        mode, lut, l0, dl = ks.choose_sqrt_pk_table(pkf, N, L, L, L, bf)
    ft = ks.filter_tables(transfer_fn, N, L, L, L)
    kmin, kmax = 2. * np.pi / L, 2. * np.pi * np.sqrt(3.) * N / L
    edges = ks.pk_bin_edges(kmin, kmax, NBINS)
    thr = ks.bin_thresholds(edges)
    tables = dict(lut=lut, tperp=ft.tperp, tpar=ft.tpar, thr=thr, lut_mode=(mode, l0, dl))
    upload_tables(plan, tables)
    return tables, edges


def upload_tables(plan, t):
    plan.set_sqrt_pk(t["lut"], *t["lut_mode"])
    plan.set_filter(t["tperp"], t["tpar"], None)
    plan.set_pk_bins(t["thr"])


def cpu_port_step(N, L, seed=11):
    """One step of the reference algorithm on the CPU (oracle port), as CosmoBox.realise_density runs it:
    the white-noise draw (box.py:174-175) and the k grid (box.py:116-127) are inside the timed region."""
    from oracle import restate as R
    from _util import pk_function
    _, pkf = pk_function(Z_BOX)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        rng = np.random.RandomState(seed)
        re = rng.normal(0., 1., (N, N, N))                                        # box.py:174
        im = rng.normal(0., 1., (N, N, N))                                        # box.py:175
        dx, dk = R.realise_density_port(re, im, pkf, N, L, L, L)                  # box.py:161-193
        filt = R.apply_transfer_fn_port(dk, transfer_fn, N, L, L, L)               # box.py:374-380
        R.binned_power_spectrum_port(np.fft.fftn(filt.real), N, L, L, L, nbins=NBINS)   # box.py:736-768
        return time.perf_counter() - t0


CPU_SAMPLE_N = 256       # one bounded box of the workload per CPU step (the reference cannot hold 1024^3)


def cpu_sample_text(n, extra=""):
    return ("%d^3 box per step: NumPy port of fastbox/box.py realise_density (incl. its np.random draws and k grid) + "
            "apply_transfer_fn + binned_power_spectrum, float64; numpy.fft is single-threaded as in the reference; "
            "the reference cannot hold 1024^3 (~140 GB of float64 temporaries); host has %d cores%s"
            % (n, os.cpu_count() or 0, extra))


def run_reference(args, rank):
    if rank != 0:
        return
    n_ref = CPU_SAMPLE_N
    L = 2000.0 * n_ref / 1024.0
    for _ in range(min(args.warmup, 1)):
        cpu_port_step(n_ref, L)
    ts = [cpu_port_step(n_ref, L, seed=11 + i) for i in range(args.steps)]
    t = float(np.mean(ts))
    val = n_ref ** 3 / t / 1e6
    sample = cpu_sample_text(n_ref)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mcells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "realise+filter+P(k), %d^3 (bounded sample of the %d^3 workload of the %d-GPU line; "
                                   "the CPU path is one process on the host cores whatever N is)"
                                   % (n_ref, args.size if args.gpus == 1 else args.size_multi, args.gpus),
                       "nbins": NBINS},
            "cpu_baseline": {"value": val, "unit": "Mcells/s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def one_gpu_sharded_workload(N, L, dev, steps=3):
    """LIVE one-GPU rate of the workload the multi-GPU lines shard (N^3, Philox noise): the strong-scaling
    denominator.  Needs 8 N^3 bytes (68 GB at 2048^3)."""
    from fastbox_b200 import _lib
    plan = _lib.Plan(N, L, L, L, dev)
    try:
        configure_plan(plan, N, L)
        field = plan.alloc(N ** 3 * 4)
        flags = _lib.F_SQRTPK | _lib.F_FILTER
        plan.realise(None, None, seed=1, flags=flags, field_out=field, want_pk=True)
        plan.sync()
        plan.timer_start()
        for it in range(steps):
            plan.realise(None, None, seed=it, flags=flags, field_out=field, want_pk=True)
        ms = plan.timer_stop() / steps
        field.free()
        return {"value": N ** 3 / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": ms, "steps": steps,
                "source": "measured live in this run on one GPU (%d^3, Philox noise, same tables)" % N}
    except Exception as exc:      # pragma: no cover
        return {"value": None, "error": str(exc)}
    finally:
        plan.close()


def run_single(args):
    import torch
    from fastbox_b200 import _lib
    N = args.size
    L = 2000.0 * N / 1024.0
    dev = 0
    torch.cuda.set_device(dev)
    plan = _lib.Plan(N, L, L, L, dev)
    tables, edges = configure_plan(plan, N, L)
    hbm_peak, peak_src = measured_peaks()
    n3 = N ** 3
    flags = _lib.F_SQRTPK | _lib.F_FILTER
    # synthetic white noise, resident in HBM (torch is only the tensor carrier)
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    re = torch.randn(n3, device="cuda", dtype=torch.float32, generator=g)
    im = torch.randn(n3, device="cuda", dtype=torch.float32, generator=g)
    field = torch.empty(n3, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()

    def step_dev():
        return plan.realise(re, im, flags=flags, field_out=field, want_pk=True, want_sums=False)

    for _ in range(args.warmup):
        step_dev()
    pass_ms = np.zeros(3)
    launches0 = _lib.launch_count()
    with ClockSampler(dev) as clk:
        plan.sync()
        plan.timer_start()
        for _ in range(args.steps):
            res, sums = step_dev()
            pass_ms += np.array(plan.last_timings(3))
        ms = plan.timer_stop()
    launches = _lib.launch_count() - launches0
    ms_step = ms / args.steps
    pass_ms /= args.steps
    value = n3 / (ms_step * 1e-3) / 1e6

    # Philox variant (no noise read), same pipeline
    for _ in range(2):
        plan.realise(None, None, seed=1, flags=flags, field_out=field, want_pk=True)
    plan.sync()
    plan.timer_start()
    for it in range(max(2, args.steps // 2)):
        plan.realise(None, None, seed=it, flags=flags, field_out=field, want_pk=True)
    ms_philox = plan.timer_stop() / max(2, args.steps // 2)

    # ---- e2e: host (pinned) buffers through the same C-ABI call
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e, tuning runs only)")
        h_re = plan.host_alloc((n3,), np.float32)
        h_im = plan.host_alloc((n3,), np.float32)
        h_field = plan.host_alloc((n3,), np.float32)
        _lib.check(plan.lib.fb_copy(plan.h, h_re.ctypes.data, re.data_ptr(), n3 * 4))
        _lib.check(plan.lib.fb_copy(plan.h, h_im.ctypes.data, im.data_ptr(), n3 * 4))
        e2e_steps = max(2, min(args.steps, 5))
        plan.realise(h_re, h_im, flags=flags, field_out=h_field, want_pk=True)       # warm staging buffers
        plan.sync()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            upload_tables(plan, tables)
            res_h, _ = plan.realise(h_re, h_im, flags=flags, field_out=h_field, want_pk=True)
        plan.sync()
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        tab_bytes = int(sum(np.asarray(v).nbytes for k, v in tables.items() if k != "lut_mode"))
        e2e = {"value": n3 / t_e2e / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": 2 * 4 * n3 + tab_bytes,
               "d2h_bytes_per_step": 4 * n3 + 3 * 8 * (NBINS + 1), "ms_per_step": t_e2e * 1e3,
               "api": "fb_realise(host re, host im -> host field, P(k) moments) via ctypes"}
        # The step is PCIe bound (12.9 GB over the bus).  Two host threads, one plan (stream + staging
        # buffers) each, let the H2D copies of one step overlap the D2H copy of the other (PCIe is full
        # duplex); every step still moves all of its inputs and outputs inside the timed region.
        try:
            import threading
            plan2 = _lib.Plan(N, L, L, L, dev)
            configure_plan(plan2, N, L)
            h_field2 = plan2.host_alloc((n3,), np.float32)
            plan2.realise(h_re, h_im, flags=flags, field_out=h_field2, want_pk=True)
            plan2.sync()
            per_thread = max(3, e2e_steps)
            errs = []

            def worker(pl, h_out):
                try:
                    for _ in range(per_thread):
                        upload_tables(pl, tables)
                        pl.realise(h_re, h_im, flags=flags, field_out=h_out, want_pk=True)
                    pl.sync()
                except Exception as exc_:      # pragma: no cover
                    errs.append(exc_)

            th = [threading.Thread(target=worker, args=(plan, h_field)),
                  threading.Thread(target=worker, args=(plan2, h_field2))]
            t0 = time.perf_counter()
            for t_ in th:
                t_.start()
            for t_ in th:
                t_.join()
            t_pipe = (time.perf_counter() - t0) / (2 * per_thread)
            if errs:
                raise errs[0]
            e2e["single_stream"] = {"value": e2e["value"], "ms_per_step": e2e["ms_per_step"]}
            if t_pipe < t_e2e:
                e2e["value"] = n3 / t_pipe / 1e6
                e2e["ms_per_step"] = t_pipe * 1e3
                e2e["api"] += "; two host threads with one plan each (H2D of one step overlaps D2H of the other)"
            else:
                e2e["two_threads"] = {"value": n3 / t_pipe / 1e6, "ms_per_step": t_pipe * 1e3}
            plan2.close()
        except Exception as exc:      # pragma: no cover
            e2e["two_threads_error"] = str(exc)
    except Exception as exc:      # pragma: no cover
        e2e = {"value": None, "unit": "Mcells/s", "error": str(exc)}

    # ---- roofline of the dominant kernel (longest pass) and of the whole step
    names = ["k_rows_inv (noise -> Hermitian spectrum, sqrtP, filter, P(k), z FFT)", "k_cols_c2c (y FFT)",
             "k_x_c2r (x FFT, half-complex -> real)"]
    nh = (N // 2 + 1) * N * N                                   # half-spectrum points (complex64)
    pass_bytes = [8.0 * n3 + 8.0 * nh, 16.0 * nh, 8.0 * nh + 4.0 * n3]     # 8+4, 4+4, 4+4 B/cell
    dom = int(np.argmax(pass_ms))
    ach = pass_bytes[dom] / (pass_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src,
                "pipeline": {"bytes_per_cell": BYTES_PER_CELL_HBM_NOISE,
                             "achieved": BYTES_PER_CELL_HBM_NOISE * n3 / (ms_step * 1e-3) / 1e9,
                             "frac": BYTES_PER_CELL_HBM_NOISE * n3 / (ms_step * 1e-3) / 1e9 / hbm_peak,
                             "frac_nominal_8TBs": BYTES_PER_CELL_HBM_NOISE * n3 / (ms_step * 1e-3) / 1e9 / 8000.0,
                             "strict_io_floor_frac": 12.0 * n3 / (ms_step * 1e-3) / 1e9 / hbm_peak},
                "pass_ms": [float(x) for x in pass_ms],
                "philox_variant": {"ms_per_step": ms_philox, "Mcells_per_s": n3 / (ms_philox * 1e-3) / 1e6,
                                   "frac": BYTES_PER_CELL_PHILOX * n3 / (ms_philox * 1e-3) / 1e9 / hbm_peak}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(["rows_inv", "cols", "x_c2r"][dom])
        except Exception:
            pass

    # ---- CPU baseline (oracle port) on a bounded sample
    cpu = None
    if not args.no_cpu:
        n_cpu = args.cpu_size
        ts_cpu = [cpu_port_step(n_cpu, 2000.0 * n_cpu / 1024.0, seed=11 + i) for i in range(2)]
        t_cpu = float(np.mean(ts_cpu))
        cpu = {"value": n_cpu ** 3 / t_cpu / 1e6, "unit": "Mcells/s", "cores": 1, "kind": "port",
               "sample": cpu_sample_text(n_cpu, "; two steps, %.1f s in total" % sum(ts_cpu))}

    # ---- the workload the multi-GPU lines shard (2048^3, Philox), measured live on this one GPU
    sharded = None
    if not args.no_one_gpu and N == 1024:
        del re, im, field
        torch.cuda.empty_cache()
        sharded = one_gpu_sharded_workload(args.size_multi, 2000.0 * args.size_multi / 1024.0, dev)

    line = {"metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%d^3 realise + k_perp/k_par filter + binned P(k) (nbins=%d), white noise "
                                   "resident in HBM" % (N, NBINS),
                       "box_Mpc": L, "redshift": Z_BOX,
                       "l2": "inputs (%.1f GB noise) larger than L2 (126 MB); no flush needed" % (8 * n3 / 1e9),
                       "sharded_workload_on_one_gpu": sharded},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu}
    print(json.dumps(line))
    plan.close()


def run_config(args):
    """
    BASELINE.json configs[1..3] as stage-by-stage device-resident pipelines (one JSON line): every stage is one
    C-ABI call on buffers resident in HBM, timed with CUDA events on the library stream; bytes/cell are the
    algorithmic figures of SURVEY section 8(d).
    """
    import torch
    from fastbox_b200 import _lib
    F = _lib
    name = args.config
    N = 512 if name == "lognormal_rsd_512" else 1024
    L = 2000.0
    dev = 0
    torch.cuda.set_device(dev)
    plan = _lib.Plan(N, L, L, L, dev)
    configure_plan(plan, N, L)
    hbm_peak, peak_src = measured_peaks()
    n3 = N ** 3
    nh = (N // 2 + 1) * N * N
    g = torch.Generator(device="cuda")
    g.manual_seed(41)
    re = torch.randn(n3, device="cuda", dtype=torch.float32, generator=g)
    im = torch.randn(n3, device="cuda", dtype=torch.float32, generator=g)
    field, f2, f3 = (plan.alloc(n3 * 4) for _ in range(3))
    spec = plan.alloc(nh * 8)
    zgrid = np.linspace(-0.5 * L, 0.5 * L, N)
    bias = 0.84081272                                    # HITracer.bias_HI(0.8) (tracers.py:129-144)
    stages = []
    extra = {}

    def stage(label, bytes_per_cell, fn):
        stages.append((label, bytes_per_cell, fn))

    if name == "lognormal_rsd_512":
        # example_endtoend.py:27-44: realise -> b delta -> log-normal -> v_z -> redshift-space remap
        stage("realise_density + store delta_k (box.py:161-193)", 32,
              lambda: plan.realise(re, im, flags=F.F_SQRTPK, field_out=field, spec_out=spec, want_sums=False))
        st = {}
        stage("bias + exp from delta_k (tracers.py:129-144, box.py:457)", 24,
              lambda: st.__setitem__("s", plan.spectrum_to_field(spec, f2, flags=F.F_EXP, scale=bias)[0]))
        stage("log-normal normalise (box.py:458-459)", 8,
              lambda: plan.affine(f2, n3, 1.0 / (st["s"] / n3), -1.0))
        stage("v_z = Re ifftn(velocity_k[2]) (box.py:254-285)", 24,
              lambda: plan.spectrum_to_field(spec, f3, kind=F.KIND_VEL_Z, scale=57.3))
        stage("redshift_space_density (box.py:412-437)", 12, lambda: plan.rsd_remap(f2, f3, None, zgrid, 100.0, field))
    elif name == "filter_beam_poles_1024":
        x = torch.arange(N, device="cuda", dtype=torch.float32) - N / 2.
        sig = (1.5 + 4.0 * torch.arange(N, device="cuda", dtype=torch.float32) / N)
        beam = torch.exp(-0.5 * (x[:, None, None] ** 2 + x[None, :, None] ** 2) / sig[None, None, :] ** 2).contiguous()
        stage("realise + k_perp/k_par filter (box.py:161-193, 374-380)", 28,
              lambda: plan.realise(re, im, flags=F.F_SQRTPK | F.F_FILTER, field_out=field, want_sums=False))
        # the beam cube of a survey is fixed: its spectrum is prepared once per plan (fb_beam_set, timed below as
        # `setup_ms`), every convolution then reads it (56 B/cell: 4+8, 8+16+8, 8+4)
        plan.beam_set(beam)
        plan.sync()
        t0 = time.perf_counter()
        plan.beam_set(beam)
        plan.sync()
        extra["beam_setup_ms"] = (time.perf_counter() - t0) * 1e3
        stage("BeamModel.convolve_fft, cached beam spectrum (beams.py:81-87)", 56, lambda: plan.beam_convolve(None, field, f2))
        stage("binned P(k) + l = 2, 4 multipoles of the observed field (box.py:736-764 ext.)", 20,
              lambda: plan.field_to_spectrum(f2, want_pk=True, poles=True))
    else:
        # example_halos.py:20-53: log-normal field -> Poisson halo counts -> cross P(k) with the 21cm field
        u = torch.rand(n3, device="cuda", dtype=torch.float64, generator=g)
        counts = plan.alloc(n3 * 4)
        nbar = np.array([1e-3], np.float32)
        b1 = np.array([1.0], np.float32)
        lam_bar = 1e-3 * L ** 3 / n3
        st = {}
        stage("realise_density + store delta_k (box.py:161-193)", 32,
              lambda: plan.realise(re, im, flags=F.F_SQRTPK, field_out=field, spec_out=spec, want_sums=False))
        stage("exp(delta) from delta_k (box.py:457)", 24,
              lambda: st.__setitem__("s", plan.spectrum_to_field(spec, f2, flags=F.F_EXP, scale=1.0)[0]))
        stage("log-normal normalise (box.py:458-459)", 8, lambda: plan.affine(f2, n3, 1.0 / (st["s"] / n3), -1.0))
        stage("halo_count_field, Poisson inversion (halos.py:91-117)", 16,
              lambda: plan.halo_counts(f2, nbar, 0, b1, 0, False, 0.0, u, counts))
        stage("halo overdensity N_h / N_bar - 1", 8, lambda: plan.counts_to_field(counts, n3, 1.0 / lam_bar, -1.0, f3))
        stage("cross P(k) of the halo field with delta_k (example_halos.py:46-53)", 24,
              lambda: st.__setitem__("pk", plan.field_to_spectrum(f3, cross=spec, want_pk=True)))

    for _ in range(max(1, args.warmup)):
        for _, _, fn in stages:
            fn()
    plan.sync()
    per = np.zeros(len(stages))
    launches0 = _lib.launch_count()
    with ClockSampler(dev) as clk:
        plan.timer_start()
        for _ in range(args.steps):
            for _, _, fn in stages:
                fn()
        total_ms = plan.timer_stop() / args.steps
        for i, (_, _, fn) in enumerate(stages):                      # per-stage times, measured separately
            plan.timer_start()
            for _ in range(3):
                fn()
            per[i] = plan.timer_stop() / 3
    launches = _lib.launch_count() - launches0
    bpc = float(sum(b for _, b, _ in stages))
    rows = [{"stage": lab, "ms": float(ms), "bytes_per_cell": b, "GBs": b * n3 / (ms * 1e-3) / 1e9,
             "frac_hbm": b * n3 / (ms * 1e-3) / 1e9 / hbm_peak} for (lab, b, _), ms in zip(stages, per)]
    dom = int(np.argmax(per))
    line = {"metric": "Mcells/s " + name, "value": n3 / (total_ms * 1e-3) / 1e6, "unit": "Mcells/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %d^3, every stage one C-ABI call on HBM-resident buffers" % (name, N),
                       "box_Mpc": L, "l2": "all cubes (%.1f GB each) larger than L2" % (4 * n3 / 1e9)},
            "stages": rows, "clocks": clk.summary(), "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": rows[dom]["stage"], "achieved": rows[dom]["GBs"], "peak": hbm_peak,
                         "unit": "GB/s", "frac": rows[dom]["frac_hbm"], "traffic": None, "peak_source": peak_src,
                         "pipeline": {"bytes_per_cell": bpc, "achieved": bpc * n3 / (total_ms * 1e-3) / 1e9,
                                      "frac": bpc * n3 / (total_ms * 1e-3) / 1e9 / hbm_peak}},
            "e2e": None, "cpu_baseline": None}
    line.update(extra)
    print(json.dumps(line))
    plan.close()


def multi_gpu_check(rank, world, local_rank, mode):
    """
    Correctness of the sharded path inside the bench run (SCALE lines carry it): a 256^3 box realised over all
    ranks equals the same box realised on one GPU -- field bit for bit, bin populations exactly, moments to 1e-12.
    """
    import torch
    import torch.distributed as dist
    from fastbox_b200 import _lib
    from fastbox_b200 import dist as fbd
    N = 256
    L = 2000.0 * N / 1024.0
    flags = _lib.F_SQRTPK | _lib.F_FILTER
    if mode == "p2p":
        dr = fbd.NvlinkRealiser(N, (L, L, L), rank, world, local_rank, chunks=2, with_forward=True)
        configure_plan(dr.plan, N, L)
        for _ in range(2):
            _, pk, _ = dr.realise(5, flags, want_pk=True)
        fwd = dr.power_spectrum()
        slab = torch.from_numpy(dr.field_host()).cuda()
        plan_close = dr.close
    else:
        eng = fbd.CudaEngine(N, (L, L, L), rank, world, local_rank, chunks=1)
        configure_plan(eng.plan, N, L)
        dn = fbd.DistributedRealiser(eng)
        slab, pk, _ = dn.realise(5, flags, want_pk=True)
        torch.cuda.synchronize()
        fwd = dn.power_spectrum()
        slab = slab.clone()
        plan_close = eng.plan.close
    parts = [torch.empty_like(slab) for _ in range(world)]
    dist.all_gather(parts, slab)
    out = None
    if rank == 0:
        full = torch.cat(parts, dim=1).cpu().numpy()
        plan = _lib.Plan(N, L, L, L, local_rank)
        configure_plan(plan, N, L)
        ref = np.empty((N, N, N), np.float32)
        res, _ = plan.realise(None, None, seed=5, flags=flags, field_out=ref, want_pk=True)
        fref = plan.field_to_spectrum(ref, want_pk=True)
        plan.close()
        nz = res["sum1"] != 0
        out = {"size": N, "field_bit_identical_to_1gpu": bool(np.array_equal(full, ref)),
               "field_rel_l2_vs_1gpu": float(np.linalg.norm(full.astype(np.float64) - ref) / np.linalg.norm(ref)),
               "pk_counts_equal": bool(np.array_equal(pk["count"], res["count"])),
               "pk_sum1_max_rel_err": float(np.max(np.abs(pk["sum1"][nz] - res["sum1"][nz]) / np.abs(res["sum1"][nz]))),
               "forward_pk_counts_equal": bool(np.array_equal(fwd["count"], fref["count"])),
               "forward_pk_sum1_max_rel_err": float(np.max(np.abs(fwd["sum1"][nz] - fref["sum1"][nz]) / np.abs(fref["sum1"][nz])))}
    plan_close()
    return out


def run_multi(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from fastbox_b200 import _lib
    from fastbox_b200 import dist as fbd
    N = args.size_multi
    L = 2000.0 * N / 1024.0
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    mode = os.environ.get("FB_DIST_MODE", "p2p")           # p2p: exchange inside the library; nccl: all_to_all_single
    chunks = int(os.environ.get("FB_CHUNKS", "16"))
    while chunks > 1 and ((N // 2) // world) % chunks:
        chunks //= 2
    flags = _lib.F_SQRTPK | _lib.F_FILTER
    hbm_peak, peak_src = measured_peaks()

    # live one-GPU rate of the very same workload (rank 0, before the sharded buffers exist)
    one_gpu = one_gpu_sharded_workload(N, L, local_rank) if (rank == 0 and not args.no_one_gpu) else None
    dist.barrier()
    check = multi_gpu_check(rank, world, local_rank, mode)
    dist.barrier()

    pass_ms = np.zeros(3)
    if mode == "p2p":
        dr = fbd.NvlinkRealiser(N, (L, L, L), rank, world, local_rank, chunks=chunks, with_forward=False)
        plan = dr.plan
        configure_plan(plan, N, L)

        def step(seed):
            return dr.realise(seed, flags, want_pk=True)
        field_dev = dr.field
    else:
        eng = fbd.CudaEngine(N, (L, L, L), rank, world, local_rank, chunks=chunks)
        plan = eng.plan
        configure_plan(plan, N, L)
        dn = fbd.DistributedRealiser(eng)

        def step(seed):
            out = dn.realise_overlapped(seed, flags, want_pk=True) if chunks > 1 else dn.realise(seed, flags, want_pk=True)
            eng.sync()
            return out
        field_dev = eng.field

    for w in range(args.warmup):
        step(w)
    launches0 = _lib.launch_count()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        # every step ends with a host synchronisation of the library stream (the P(k) read-back), so events
        # on torch's stream bracket all device work of the steps
        ev0.record()
        for s_ in range(args.steps):
            _, pk, _ = step(s_)
            if mode == "p2p":
                pass_ms += np.array(plan.last_timings(3))
        ev1.record()
        torch.cuda.synchronize()
    t_wall = ev0.elapsed_time(ev1) * 1e-3
    launches = _lib.launch_count() - launches0
    pass_ms /= args.steps

    # tuning / evidence: other exchange mechanisms measured in the same run (FB_DIST_COMPARE="1,3,2:8" = xmode or
    # xmode:push CTAs per peer), same steps, max over ranks; each is checked against the default's field moments
    alternatives = None
    if mode == "p2p" and os.environ.get("FB_DIST_COMPARE"):
        alternatives = []
        base_mode = int(os.environ.get("FB_DIST_XMODE", str(DIST_XMODE_DEFAULT)))
        default_push_ctas = int(os.environ.get("FB_DIST_PUSH_CTAS", "0")) or max(4, 32 // max(1, world - 1))
        _, _, sums0 = dr.realise(0, flags, want_pk=True, want_sums=True)
        for spec in os.environ["FB_DIST_COMPARE"].split(","):
            xm, _, he = spec.partition(":")
            plan.dist_set_option("xmode", int(xm))
            plan.dist_set_option("push_ctas", int(he) if he else default_push_ctas)
            _, pk_a, sums_a = dr.realise(0, flags, want_pk=True, want_sums=True)
            step(1)
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for s_ in range(args.steps):
                step(s_)
            b.record()
            torch.cuda.synchronize()
            tt = torch.tensor([a.elapsed_time(b) / args.steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            same = abs(sums_a[0] - sums0[0]) <= 1e-9 * abs(sums0[1]) ** 0.5 + 1e-9 * abs(sums0[0]) and \
                abs(sums_a[1] - sums0[1]) <= 1e-10 * abs(sums0[1])
            alternatives.append({"xmode": int(xm), "push_ctas": int(he) if he else None, "ms_per_step": float(tt[0]),
                                 "field_moments_equal_default": bool(same),
                                 "pk_count_sum_ok": bool(int(pk_a["count"].sum()) == N ** 3)})
        plan.dist_set_option("xmode", base_mode)
        plan.dist_set_option("push_ctas", default_push_ctas)

    # result checks on the full-size box itself: every mode binned once, Parseval (box.py:944-946)
    if mode == "p2p":
        _, pk, sums = dr.realise(0, flags, want_pk=True, want_sums=True)
        sumsq = torch.tensor([sums[1]], device="cuda", dtype=torch.float64)
    else:
        out = step(0)
        pk = out[1]
        sumsq = torch.tensor([out[2][1]], device="cuda", dtype=torch.float64)
    dist.all_reduce(sumsq)
    parseval = float(sumsq[0]) * N ** 3 / (float(pk["sum1"].sum()) * (N ** 6. / L ** 3))

    # the exchange alone, for the NVLink roofline
    a2a = max(fbd.alltoall_bytes_per_rank(N, world))
    push_sweep = None
    if mode == "p2p":
        dist.barrier()
        t_x_alone = plan.dist_bench_exchange(3) * 1e-3
        if os.environ.get("FB_PUSH_SWEEP") and int(os.environ.get("FB_DIST_XMODE", str(DIST_XMODE_DEFAULT))) >= 2:
            # tuning: the copy kernel alone for several CTA counts per peer (collective: same order on all ranks)
            push_sweep = {}
            keep = int(os.environ.get("FB_DIST_PUSH_CTAS", "0"))
            for c in (2, 3, 4, 6, 8, 12, 16, 24):
                plan.dist_set_option("push_ctas", c)
                dist.barrier()
                tt = torch.tensor([plan.dist_bench_exchange(3)], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                push_sweep[str(c)] = {"ms": float(tt[0]), "GBs": a2a / (float(tt[0]) * 1e-3) / 1e9}
            plan.dist_set_option("push_ctas", keep if keep > 0 else max(4, 32 // max(1, world - 1)))
    else:
        xs = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            a.record()
            dn.exchange()
            b.record()
            torch.cuda.synchronize()
            xs.append(a.elapsed_time(b) * 1e-3)
        t_x_alone = min(xs)

    # end to end: every step also brings the rank's field slab to pinned host memory
    ny = N // world
    host_slab = plan.host_alloc((N, ny, N), np.float32)
    e2e_steps = max(2, min(args.steps, 4))
    step(0)
    _lib.check(plan.lib.fb_copy(plan.h, host_slab.ctypes.data, _lib._ptr(field_dev), host_slab.nbytes))
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s_ in range(e2e_steps):
        step(s_)
        _lib.check(plan.lib.fb_copy(plan.h, host_slab.ctypes.data, _lib._ptr(field_dev), host_slab.nbytes))
    torch.cuda.synchronize()
    dist.barrier()
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([t_wall, t_x_alone, t_e2e] + list(pass_ms), device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_wall, t_x_alone, t_e2e = float(t[0]), float(t[1]), float(t[2])
    pass_ms = [float(x) for x in t[3:6]]
    if rank == 0:
        ms_step = t_wall / args.steps * 1e3
        value = N ** 3 / (ms_step * 1e-3) / 1e6
        nv = a2a / t_x_alone / 1e9
        check = dict(check or {}, count_sum=int(pk["count"].sum()), count_sum_expected=N ** 3,
                     count_sum_ok=bool(int(pk["count"].sum()) == N ** 3), parseval_ratio=parseval)
        xmode = int(os.environ.get("FB_DIST_XMODE", str(DIST_XMODE_DEFAULT)))
        mech = {0: "the y pass stores straight into the peers' receive buffers",
                1: "the y pass writes per-destination blocks that the copy engines push into the peers' receive buffers",
                2: "the y pass writes per-destination blocks that a high-priority copy kernel (a few CTAs per peer, "
                   "16-byte peer stores) pushes into the peers' receive buffers while the k-space passes of the next "
                   "chunk run",
                3: "the y pass writes per-destination blocks that a high-priority copy kernel pushes into the peers' "
                   "receive buffers through the bulk copy engine (cp.async.bulk, one thread per CTA) while the "
                   "k-space passes of the next chunk run"}[xmode]
        how = ("exchange inside the library over NVLink peer memory (CUDA IPC): %s; %d chunks of planes, device-side "
               "epoch flags, no NCCL on the data path" % (mech, chunks)) if mode == "p2p" else \
              ("one NCCL all_to_all_single in %d chunks overlapped with the k-space passes" % chunks)
        nvlink = {"alternatives": alternatives, "bytes_sent_per_gpu": a2a, "exchange_alone_ms": t_x_alone * 1e3, "achieved_GBs": nv, "push_sweep": push_sweep,
                  "frac_of_900": nv / 900.0, "frac_of_measured_770": nv / 770.0, "mode": mode}
        if mode == "p2p":
            nvlink.update({"in_pipeline": {"kspace_passes_with_peer_stores_ms": pass_ms[0], "wait_for_peers_ms": pass_ms[1],
                                           "x_pass_ms": pass_ms[2],
                                           "note": "max over ranks of the per-step means; the first segment contains "
                                                   "the first pass, the y pass and its NVLink stores"},
                           "in_pipeline_GBs": a2a / (pass_ms[0] * 1e-3) / 1e9 if pass_ms[0] > 0 else None})
        line = {"metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "%d^3 realise + filter + binned P(k), Philox noise, slab decomposition over "
                                       "%d GPUs; %s" % (N, world, how),
                           "box_Mpc": L,
                           "l2": "per-GPU working set %.1f GB >> L2" % (12.0 * N ** 3 / world / 1e9),
                           "same_workload_on_one_gpu": one_gpu},
                "check": check,
                "clocks": clk.summary(), "gpu_launches": int(launches),
                "e2e": {"value": N ** 3 / t_e2e / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": 8,
                        "d2h_bytes_per_step": 4 * N ** 3 + world * 3 * 8 * (NBINS + 1), "ms_per_step": t_e2e * 1e3,
                        "note": "Philox noise is generated on the device (the 8-byte seed is the only input); every "
                                "step each rank copies its field slab to pinned host memory and reads back the P(k) "
                                "moments"},
                "roofline": {"bound": "hbm", "achieved": BYTES_PER_CELL_PHILOX * N ** 3 / world / (ms_step * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": BYTES_PER_CELL_PHILOX * N ** 3 / world / (ms_step * 1e-3) / 1e9 / hbm_peak,
                             "traffic": None, "peak_source": peak_src, "nvlink": nvlink},
                "cpu_baseline": None}
        if one_gpu and one_gpu.get("value"):
            line["strong_scaling_vs_live_one_gpu"] = {"speedup": value / one_gpu["value"],
                                                      "efficiency": value / one_gpu["value"] / world}
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="grid size for the single-GPU run")
    ap.add_argument("--size-multi", type=int, default=2048, help="grid size for the multi-GPU run")
    ap.add_argument("--cpu-size", type=int, default=CPU_SAMPLE_N)
    ap.add_argument("--no-one-gpu", action="store_true", help="skip the live one-GPU run of the sharded workload")
    ap.add_argument("--config", default="headline",
                    choices=["headline", "lognormal_rsd_512", "filter_beam_poles_1024", "halos_cross_1024"],
                    help="BASELINE.json configs[1..3] as stage-by-stage pipelines (one JSON line each)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs: skip the host-buffer leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        run_multi(args, rank, world, local_rank)
    elif args.config != "headline":
        run_config(args)
    else:
        run_single(args)


if __name__ == "__main__":
    main()
