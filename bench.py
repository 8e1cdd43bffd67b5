#!/usr/bin/env python
"""bench.py — headline benchmark of the two SynthObs hot paths on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  Primary record: BASELINE.json config[1]
(SED_fluxes-style: galaxies of 1e5 particles, BPASS-shape grid 51x13x9000, ISM+BC LOS
dust, 10 broadband filters), batched so one step streams more than L2 from HBM.
`value` = galaxy SEDs x filters / s with inputs resident in HBM; `e2e` = the same through
the public API with pinned HOST inputs (H2D + D2H inside the timed region).
Secondary record under "imaging": BASELINE.json config[2] (1e6 particles, 8 filters,
512x512 observed-frame images, super-sampling 2, adaptive smoothing, PSF convolution).
Multi-GPU: one process per GPU, galaxies/objects sharded by rank, no data-path collective
(weak scaling); time = max over ranks.

--impl reference times the CPU restatement of the reference (oracle/, NumPy, the reference is
pure Python so there is nothing to compile into oracle/_ref) on bounded samples of the same
workload on the host cores.
"""

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PER_GALAXY = 100000
N_FILTERS = 10
N_LAM = 9000
GALAXIES_PER_GPU = 256          # 2.56e7 particles, 1.02 GB of float64 inputs per step (>> 126 MB L2)
IMG_PARTICLES = 1000000
IMG_FILTERS = 8
IMG_NDIM = 512
IMG_SS = 2
IMG_K = 8
Z = 8.0


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            p = json.load(fh)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, 20 ms period;
    same fields as the nvidia-smi clocks line of B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.max_mhz = 0.0

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = int(vis.split(',')[self.index]) if vis and vis.split(',')[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:   # noqa: BLE001
            self.nv = None
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = get_reasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.time(), float(clk), int(r), pw))
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.005)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1)

    def summary(self, t0, t1):
        if self.nv is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml unavailable: ' + getattr(self, 'err', '')]}
        # reason bit masks (nvml.h): sw_power_cap 0x4, hw_slowdown 0x8, sw_thermal 0x20, hw_thermal 0x40
        names = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown'}
        sel = [x for x in self.samples if t0 <= x[0] <= t1] or self.samples[-3:]
        busy = [x for x in sel if x[3] > 200.0]          # samples taken while the GPU was actually under load
        sel = busy or sel
        reasons = set()
        for _, _, r, _ in sel:
            for bit, name in names.items():
                if r & bit:
                    reasons.add(name)
        return {'sm_mhz': float(np.median([x[1] for x in sel])) if sel else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(reasons), 'samples': len(sel),
                'power_w_max': max([x[3] for x in sel]) if sel else None}


def timed_steps(step, steps, barrier, flush=None):
    """Mean device time of `steps` runs of `step`, each bracketed by its own CUDA events; when `flush` (a buffer
    larger than L2) is given it is rewritten between the runs, outside the timed intervals."""
    import torch
    barrier()
    tot = 0.0
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    barrier()
    return tot / steps


def setup_model(n_lam=N_LAM, n_filters=N_FILTERS):
    from synthobs_b200 import synthetic
    from synthobs_b200.sed import models
    grid = synthetic.ssp_grid(n_lam=n_lam, seed=0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = models.define_model(grid)
    model.dust_ISM = ('simple', {'slope': -1.0})
    model.dust_BC = ('simple', {'slope': -1.0})
    F = synthetic.filter_set(model.lam * (1. + Z), n_filters=n_filters, kind='gauss', lo=12000., hi=50000.,
                             prefix='JWST.FAKE.')
    cosmo = synthetic.FixedCosmology()
    model.create_Fnu_grid(F, Z, cosmo)
    return model, F, cosmo


# ------------------------------------------------------------------ ours (GPU)

def run_ours(args):
    import torch
    import torch.distributed as dist
    from synthobs_b200 import _cabi, synthetic
    from synthobs_b200.morph import images
    from synthobs_b200.sed import models

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: point fd 1 at stderr until
        # the first collective has run, so that stdout carries the one JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
            t = torch.zeros(1, device=torch.device('cuda', local))
            dist.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _cabi.require_device()
    dev = torch.device('cuda', local)
    hbm_peak, peak_src = peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- SED workload (primary)
    model, F, cosmo = setup_model()
    G = GALAXIES_PER_GPU
    batch = synthetic.galaxy_batch(G, fixed_n=N_PER_GALAXY, seed=100 + rank)
    keys = ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC')
    host = {k: torch.from_numpy(batch[k]).pin_memory() for k in keys}
    offsets = torch.from_numpy(batch['offsets'])
    d = {k: host[k].to(dev) for k in keys}
    d_off = offsets.to(dev)
    n_part = d['Masses'].numel()
    kw_names = dict(fesc=0.0, log10t_BC=7.0)

    def sed_step_device():
        return models.generate_Fnu_batch(model, F, d['Masses'], d['Ages'], d['Metallicities'], d_off,
                                         tauVs_ISM=d['tauVs_ISM'], tauVs_BC=d['tauVs_BC'], **kw_names)

    def sed_step_e2e():
        # public API with HOST (pinned) inputs: H2D of the five particle arrays + kernels + D2H of [G, F]
        out = models.generate_Fnu_batch(model, F, host['Masses'], host['Ages'], host['Metallicities'], offsets,
                                        tauVs_ISM=host['tauVs_ISM'], tauVs_BC=host['tauVs_BC'], **kw_names)
        return out

    sampler = ClockSampler(local)
    sampler.start()
    sed_steps = 1 if args.no_sed else args.steps
    for _ in range(1 if args.no_sed else max(args.warmup, 3)):
        out = sed_step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(sed_steps):
        out = sed_step_device()
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / sed_steps)
    sed_value = world * G * N_FILTERS / (ms_step * 1e-3)
    algo_bytes = 40.0 * n_part                     # 5 float64 arrays read once
    achieved = algo_bytes / (ms_step * 1e-3) / 1e9

    # e2e
    e2e_steps = 1 if args.no_sed else max(3, min(args.steps, 10))
    for _ in range(0 if args.no_sed else 2):
        sed_step_e2e()
    barrier()
    t0 = time.time()
    for _ in range(e2e_steps):
        res = sed_step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.time() - t0) * 1e3 / e2e_steps)
    e2e_value = world * G * N_FILTERS / (e2e_ms * 1e-3)
    h2d = int(sum(host[k].numel() * 8 for k in keys) + offsets.numel() * 8)
    d2h = int(G * N_FILTERS * 8)

    # ---------------- full spectra (generate_SED) + the two tensor-core contractions (secondary)
    spectra = None
    if not args.no_spectra:
        spectra = run_spectra(args, model, F, rank, world, dev, barrier, max_over_ranks)

    # ---------------- imaging workload (secondary, config[2])
    imaging = None
    if not args.no_imaging:
        imaging = run_imaging(args, rank, world, dev, barrier, max_over_ranks, hbm_peak)
    sampler.stop()
    clocks = sampler.summary(t_wall0, time.time())     # all timed regions of this run (SED, e2e, spectra, imaging)

    # ---------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_sed(batch, model, F, budget_s=12.0)
        if imaging is not None:
            imaging['cpu_baseline'] = cpu_baseline_imaging(budget_s=12.0)

    if rank == 0:
        rec = {
            'metric': 'galaxy SEDs x filters / s', 'value': sed_value, 'unit': 'galaxy_SEDs_x_filters/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 values, f64 index compares',
            'data': 'synthetic',
            'config': {'workload': 'config[1] SED_fluxes: %d galaxies/GPU x 1e5 particles, BPASS-shape grid 51x13x%d, '
                                   'ISM+BC LOS dust, %d filters; one segmented so_sed_broadband pass per step' %
                                   (G, N_LAM, N_FILTERS),
                       'particles_per_step_per_gpu': n_part, 'l2_policy': 'inputs (1.02 GB/step) larger than L2',
                       'parallelism': 'galaxies sharded by rank, no collective'},
            'particles_per_s': world * n_part / (ms_step * 1e-3),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
                         'frac': achieved / hbm_peak, 'traffic': 1.0397e9 if n_part == 25600000 else None,
                         'traffic_source': 'ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch '
                                           '(profiles/r01_ncu_kernels.md): 1.031 GB read + 8.3 MB written vs 1.024 GB algorithmic',
                         'peak_source': peak_src,
                         'kernel': 'sed_broadband_kernel<12,0>', 'algorithmic_bytes_per_particle': 40,
                         'note': 'duration = whole step (3 launches chained by programmatic dependent launch: item scan, '
                                 'broadband, per-galaxy reduce), so the fraction is a lower bound for the '
                                 'streaming kernel'},
            'e2e': {'value': e2e_value, 'unit': 'galaxy_SEDs_x_filters/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': e2e_ms},
            'gpu_launches': 3 * sed_steps,
            'clocks': clocks,
            'cpu_baseline': cpu,
            'spectra': spectra,
            'imaging': imaging,
        }
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


def run_spectra(args, model, F, rank, world, dev, barrier, max_over_ranks):
    """generate_SED for a batch of galaxies (five spectra over 9000 wavelengths each: segmented
    histogram -> 3xTF32 tcgen05 GEMM against the SSP grid, per-particle dust kernel) followed by the
    second contraction spectra x filter weights; plus the histogram x grid GEMM alone at FLARES batch
    size (1e4 galaxies -> M = 2e4 rows) with its tensor-pipe roofline."""
    import torch
    from synthobs_b200 import _cabi, synthetic
    from synthobs_b200.sed import models
    G = 32
    batch = synthetic.galaxy_batch(G, fixed_n=N_PER_GALAXY, seed=300 + rank)
    keys = ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC')
    d = {k: torch.from_numpy(batch[k]).to(dev) for k in keys}
    off = torch.from_numpy(batch['offsets']).to(dev)
    cosmo = synthetic.FixedCosmology()

    def step():
        o = models.sed_spectra_device(model, d['Masses'], d['Ages'], d['Metallicities'], d['tauVs_ISM'], d['tauVs_BC'],
                                      0.0, 7.0, offsets=off)
        return models.spectra_Fnu(model, F, o['total'], Z, cosmo)      # [G, F] band fluxes from the spectra

    for _ in range(2):
        step()
    barrier()
    steps = max(3, min(args.steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # 256 MB > 126 MB of L2
    ms = max_over_ranks(timed_steps(step, steps, barrier, flush))
    n_lam = model.lam.size
    n_part = G * N_PER_GALAXY
    # the hist x grid GEMM alone, FLARES batch size
    ops = models._spectral_operands(model)
    M = 20000
    A = torch.rand((M, ops['ld']), device=dev)
    Ah, Al = models.split_tf32(A)
    Bh, Bl = ops['BT_split']
    C = torch.empty((M, n_lam), dtype=torch.float32, device=dev)
    K = 2 * model.na * model.nz

    def gemm_only():
        _cabi.check(_cabi.lib.so_gemm_3xtf32(_cabi.ptr(Ah), _cabi.ptr(Al), A.stride(0), _cabi.ptr(Bh), _cabi.ptr(Bl),
                                             Bh.stride(0), _cabi.ptr(C), C.stride(0), M, n_lam, K, _cabi.stream()),
                    'so_gemm_3xtf32')
    for _ in range(3):
        gemm_only()
    barrier()
    e0.record()
    for _ in range(10):
        gemm_only()
    e1.record()
    barrier()
    gms = max_over_ranks(e0.elapsed_time(e1) / 10)
    flops = 2.0 * M * n_lam * K
    try:
        bf16_peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['bf16_tflops'])
        src = 'measured bf16 burst / 2 (no TF32 entry in MEASURED_PEAKS.json)'
    except Exception:
        bf16_peak, src = 1590.0, 'fallback bf16 / 2'
    tf32_peak = bf16_peak / 2
    # context for the denominator: cuBLAS TF32 (one pass, no error compensation) on 8192^3, measured live
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    a32 = torch.rand((8192, 8192), device=dev)
    b32 = torch.rand((8192, 8192), device=dev)
    for _ in range(3):
        c32 = a32 @ b32
    barrier()
    e0.record()
    for _ in range(10):
        c32 = a32 @ b32
    e1.record()
    barrier()
    cublas_tf32 = 2.0 * 8192 ** 3 / (e0.elapsed_time(e1) / 10) / 1e9
    torch.backends.cuda.matmul.allow_tf32 = old_tf32
    del a32, b32, c32
    return {
        'metric': 'galaxy SEDs/sec (5 spectra x %d wavelengths, ISM+BC dust) + %d band fluxes' % (n_lam, len(F['filters'])),
        'value': world * G / (ms * 1e-3), 'unit': 'galaxy_SEDs/s', 'ms_per_step': ms,
        'galaxy_SEDs_x_filters_per_s': world * G * len(F['filters']) / (ms * 1e-3),
        'particle_wavelengths_per_s': world * n_part * n_lam / (ms * 1e-3),
        'config': {'workload': '%d galaxies x 1e5 particles per GPU, generate_SED_batch + spectra x filters' % G,
                   'l2_policy': 'L2 flushed (256 MB write) between steps, outside the timed intervals'},
        'ex2_roofline': _ex2_roofline(batch, n_part, n_lam, ms),
        'gemm_hist_x_grid': {'M': M, 'N': n_lam, 'K': K, 'ms': gms,
                             'l2_policy': 'operands and result (1.03 GB per launch) larger than L2', 'algorithmic_tflops': flops / gms / 1e9,
                             'executed_tf32_tflops': 3 * flops / gms / 1e9,
                             'roofline': {'bound': 'tensor', 'achieved': 3 * flops / gms / 1e9, 'peak': tf32_peak,
                                          'unit': 'TFLOP/s', 'frac': 3 * flops / gms / 1e9 / tf32_peak,
                                          'traffic': 1.96e9 if (M, n_lam, K) == (20000, 9000, 1326) else None,
                                          'traffic_note': 'ncu dram read 1.26 GB + write 0.70 GB per launch vs 1.03 GB of '
                                                          'operands (hi+lo) and result (profiles/r01_ncu_kernels.md)',
                                          'peak_source': src,
                                          'cublas_tf32_8192_tflops': cublas_tf32,
                                          'note': 'achieved counts the three TF32 MMAs actually executed per '
                                                  'algorithmic product (3xTF32); algorithmic fraction is a third'}},
    }


def imaging_stage_times(obs, Xd, Yd, Ld):
    """Device time of the stages of one imaging step (CUDA events between the stage calls; the stages are
    the same C-ABI calls particle_device_multi makes)."""
    import torch
    from synthobs_b200.morph import images
    o0 = obs[0]
    spec = images.KernelSpectrum(np.stack([o.psf_kernel() for o in obs]), o0.ndim_super, o0.ndim_super)

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e
    names = ['recentre (two medians)', 'nearest-pixel index', 'k-NN', 'plan + binning + deposit', 'PSF FFT', 'rebin']
    acc = np.zeros(len(names))
    reps = 5
    for it in range(reps + 1):
        t = [ev()]
        mx, my = images.median_device_pair(Xd, Yd)
        X, Y = Xd - mx, Yd - my
        t.append(ev())
        width, G, Gd = images.pixel_grid_device(o0.resolution, o0.ndim_super)
        ix, iy = images.ngp_index_device(X, Y, Gd, o0.ndim_super, width / 2.)
        t.append(ev())
        dk = images.knn_device(X, Y, ix, IMG_K, width / 2., floor=images.FWHM_FLOOR)
        t.append(ev())
        r = images.deposit_device(X, Y, Ld, o0.resolution, o0.ndim_super, adaptive_k=IMG_K, want_ngp=False, dk=dk,
                                  pad_to=(spec.py, spec.px))
        t.append(ev())
        conv = images.convolve_fft_device(r['data'], spec, padded=r['data_padded'], crop_copy=False)
        t.append(ev())
        images.rebin_device(r['data'], o0.ndim)
        images.rebin_device(conv, o0.ndim)
        t.append(ev())
        torch.cuda.synchronize()
        if it:
            acc += [t[i].elapsed_time(t[i + 1]) for i in range(len(names))]
    return {n: float(v / reps) for n, v in zip(names, acc)}


def _ex2_roofline(batch, n_part, n_lam, ms):
    """The dust-spectra kernel is bound by exponentials: one per (old particle, wavelength), two per young one
    (models.py:285-304).  Peak = 16 MUFU.EX2 per clock and SM at the maximum SM clock; the step time (which also
    holds the histogram, sort and GEMM launches) is used, so the fraction is a lower bound for the kernel.  The
    kernel evaluates a quarter of them on the FMA pipe instead, which is how it can pass the MUFU-only roof."""
    young = float((batch['Ages'] < 10.0).mean())
    exps = n_part * n_lam * (1.0 + young)
    peak = 148 * 16 * 1.965e9
    return {'bound': 'mufu_ex2', 'achieved': exps / (ms * 1e-3), 'peak': peak, 'unit': 'ex2/s',
            'frac': exps / (ms * 1e-3) / peak, 'young_fraction': young,
            'peak_source': '148 SMs x 16 MUFU.EX2/clk x 1.965 GHz (nominal XU rate; no measured entry)'}


def run_imaging(args, rank, world, dev, barrier, max_over_ranks, hbm_peak):
    import torch
    from synthobs_b200 import synthetic
    from synthobs_b200.morph import images
    p = synthetic.particles(IMG_PARTICLES, seed=200 + rank, scale_kpc=3.0)
    filters = ['JWST.NIRCAM.B%d' % i for i in range(IMG_FILTERS)]
    for f in filters:
        images.filter_info[f] = {'pixel_scale': 0.031}
    cosmo = synthetic.FixedCosmology()
    width_arcsec = IMG_NDIM * 0.031 + 1e-9
    rng = np.random.default_rng(7 + rank)
    Lh = np.stack([p['Masses'] * (0.5 + rng.random(IMG_PARTICLES)) for _ in filters])
    PSFs = {f: synthetic.GaussianPSF(1.5 + 0.2 * i) for i, f in enumerate(filters)}
    Xh = torch.from_numpy(p['X']).pin_memory()
    Yh = torch.from_numpy(p['Y']).pin_memory()
    Lh_t = torch.from_numpy(Lh).pin_memory()
    Xd, Yd, Ld = Xh.to(dev), Yh.to(dev), Lh_t.to(dev)
    obs = [images.observed(f, cosmo, Z, width_arcsec, smoothing=('adaptive', float(IMG_K)), PSF=PSFs[f],
                           super_sampling=IMG_SS) for f in filters]
    assert obs[0].ndim == IMG_NDIM, obs[0].ndim
    stats = {}

    def step_device(Xd, Yd, Ld):
        mx, my = images.median_device_pair(Xd, Yd)    # recentring (images.py:183-184, np.median semantics)
        X = Xd - mx                                   # on device copies
        Y = Yd - my
        r = images.particle_device_multi(obs, X, Y, Ld)
        stats.update(r['stats'])
        return r['data']

    def step_e2e():
        # public host-input entry point: pinned X, Y, L in, final images out (H2D of L overlaps the k-NN)
        out, _ = images.particle_host_multi(obs, Xh, Yh, Lh_t)
        return out

    for _ in range(3):
        step_device(Xd, Yd, Ld)
    barrier()
    steps = max(3, min(args.steps, 10))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # inputs (80 MB) fit in L2: flush between steps
    ms = max_over_ranks(timed_steps(lambda: step_device(Xd, Yd, Ld), steps, barrier, flush))
    del flush
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.time()
    for _ in range(steps):
        step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.time() - t0) * 1e3 / steps)
    stages = imaging_stage_times(obs, Xd, Yd, Ld)
    deposits = stats.get('deposits', 0.0) * IMG_FILTERS
    nsel = stats.get('selected', 0.0)
    ref_equiv = nsel * (IMG_NDIM * IMG_SS) ** 2 * IMG_FILTERS
    # algorithmic bytes of one step: positions and weights as the API hands them over (float64),
    # every super-sampled and final pixel written once
    algo_bytes = IMG_PARTICLES * (16 + 8 * IMG_FILTERS) + 4 * IMG_FILTERS * ((IMG_NDIM * IMG_SS) ** 2 + IMG_NDIM ** 2)
    ach = algo_bytes / (ms * 1e-3) / 1e9
    return {
        'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                     'traffic': None, 'algorithmic_bytes_per_step': algo_bytes,
                     'note': 'whole imaging step (recentring, k-NN, tile binning, deposit, PSF FFT, rebin: ~60 '
                             'launches), not one kernel: at ~18 pixels per particle the step is a chain of '
                             'latency-bound stages (tree walk, radix sorts, FFT) and sits far below the HBM roof; '
                             'per-kernel figures are in profiles/'},
        'stages_ms': stages,
        'metric': 'particles/sec imaged', 'unit': 'particles/s (all %d filter images)' % IMG_FILTERS,
        'value': world * IMG_PARTICLES / (ms * 1e-3), 'ms_per_step': ms,
        'particle_filter_images_per_s': world * IMG_PARTICLES * IMG_FILTERS / (ms * 1e-3),
        'deposits_evaluated_per_s': world * deposits / (ms * 1e-3),
        'reference_equivalent_deposits_per_s': world * ref_equiv / (ms * 1e-3),
        'config': {'workload': 'config[2] observed-frame: %d particles, %d filters, %dx%d, super_sampling %d, '
                               "('adaptive', %d), Gaussian PSF FFT convolution, rebin" %
                               (IMG_PARTICLES, IMG_FILTERS, IMG_NDIM, IMG_NDIM, IMG_SS, IMG_K),
                   'l2_policy': 'L2 flushed (256 MB write) between steps, outside the timed intervals',
                   'support_sigmas': images.SUPPORT_SIGMAS, 'pairs_particle_tile': stats.get('pairs'),
                   'selected': nsel, 'deposits_per_filter': stats.get('deposits')},
        'e2e': {'value': world * IMG_PARTICLES / (e2e_ms * 1e-3), 'unit': 'particles/s', 'ms_per_step': e2e_ms,
                'h2d_bytes_per_step': int(IMG_PARTICLES * 8 * (2 + IMG_FILTERS)),
                'd2h_bytes_per_step': int(IMG_FILTERS * IMG_NDIM * IMG_NDIM * 4)},
    }


# ------------------------------------------------------------------ CPU baselines (oracle port)

def cpu_baseline_sed(batch, model, F, budget_s=12.0):
    from oracle import sed_oracle
    from synthobs_b200 import providers
    tabs = {f: {c: getattr(model.Fnu[f], c) for c in sed_oracle.COMPONENTS} for f in F['filters']}
    grid = model.grid
    k = {f: float(providers.simple({'slope': -1.0}).tau(F[f].pivwv() / (1. + Z))) for f in F['filters']}
    off = batch['offsets']
    done, t0 = 0, time.time()
    while time.time() - t0 < budget_s and done < len(off) - 1:
        s, e = off[done], off[done + 1]
        sed_oracle.broadband(tabs, F['filters'], grid['log10age'], grid['log10Z'], batch['Masses'][s:e],
                             batch['Ages'][s:e], batch['Metallicities'][s:e], tauVs_ISM=batch['tauVs_ISM'][s:e],
                             tauVs_BC=batch['tauVs_BC'][s:e], k_ISM=k, k_BC=k, fesc=0.0)
        done += 1
    dt = time.time() - t0
    return {'value': done * len(F['filters']) / dt, 'unit': 'galaxy_SEDs_x_filters/s', 'cores': 1, 'kind': 'port',
            'sample': '%d galaxies x 1e5 particles x %d filters with oracle/sed_oracle.broadband (vectorised '
                      'NumPy restatement of generate_Fnu; the reference itself is a per-particle Python loop, '
                      '~1e5 particle-filters/s) in %.1f s; host has %d cores' % (done, len(F['filters']), dt,
                                                                               os.cpu_count())}


def cpu_baseline_imaging(budget_s=12.0, n_sample=4000):
    from oracle import image_oracle
    from synthobs_b200 import synthetic
    p = synthetic.particles(IMG_PARTICLES, seed=200, scale_kpc=3.0)
    cosmo = synthetic.FixedCosmology()
    g = image_oracle.observed_geometry(0.031, cosmo.arcsec_per_kpc, IMG_NDIM * 0.031 + 1e-9, super_sampling=IMG_SS)
    width, G = image_oracle.pixel_grid(g['resolution'], g['ndim_super'])
    X, Y = p['X'] - np.median(p['X']), p['Y'] - np.median(p['Y'])
    t0 = time.time()
    dk = image_oracle.kth_neighbour_distance(X, Y, IMG_K)
    t_knn = time.time() - t0
    idx = np.random.default_rng(0).choice(IMG_PARTICLES, n_sample, replace=False)
    t0 = time.time()
    done = 0
    for s in range(0, n_sample, 500):
        sl = idx[s:s + 500]
        image_oracle.adaptive_separable(G, X[sl], Y[sl], p['Masses'][sl], dk[sl], chunk=500)
        done += sl.size
        if time.time() - t0 > budget_s:
            break
    t_dep = time.time() - t0
    per_particle = t_dep / done
    total = t_knn + per_particle * IMG_PARTICLES * IMG_FILTERS
    return {'value': IMG_PARTICLES / total, 'unit': 'particles/s (all %d filter images)' % IMG_FILTERS, 'cores': 1,
            'kind': 'port',
            'sample': 'cKDTree k=%d on all 1e6 particles (%.2f s, all cores) + oracle adaptive_separable (NumPy/BLAS '
                      'restatement, full-image support like the reference) on %d particles at %dx%d: %.3f ms/particle, '
                      'extrapolated x%d particles x %d filters; PSF FFT and rebin not included' %
                      (IMG_K, t_knn, done, g['ndim_super'], g['ndim_super'], per_particle * 1e3, IMG_PARTICLES,
                       IMG_FILTERS)}


# ------------------------------------------------------------------ reference arm

_REF = {}


def _ref_one_galaxy(gi):
    from oracle import sed_oracle
    b, off = _REF['batch'], _REF['batch']['offsets']
    s, e = off[gi], off[gi + 1]
    out = sed_oracle.broadband(_REF['tabs'], _REF['filters'], _REF['grid']['log10age'], _REF['grid']['log10Z'],
                               b['Masses'][s:e], b['Ages'][s:e], b['Metallicities'][s:e], tauVs_ISM=b['tauVs_ISM'][s:e],
                               tauVs_BC=b['tauVs_BC'][s:e], k_ISM=_REF['k'], k_BC=_REF['k'], fesc=0.0)
    return [out[f] for f in _REF['filters']]


def run_reference(args):
    """The reference's own CPU path for the headline metric, on the host cores of the box: the oracle port
    (vectorised NumPy restatement of generate_Fnu; the reference is pure Python, so there is no oracle/_ref
    to compile), one galaxy per task over a pool of worker processes (galaxies are independent)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', '1'))
    import multiprocessing as mp
    from oracle import sed_oracle
    from synthobs_b200 import providers, synthetic
    grid = synthetic.ssp_grid(n_lam=N_LAM, seed=0)
    sed_oracle.ensure_incident(grid)
    F = synthetic.filter_set(grid['lam'] * (1. + Z), n_filters=N_FILTERS, kind='gauss', lo=12000., hi=50000.,
                             prefix='JWST.FAKE.')
    cosmo = synthetic.FixedCosmology()
    t0 = time.time()
    tabs = sed_oracle.fnu_grid(grid, F, Z, cosmo.dl_cm, providers.madau)
    t_setup = time.time() - t0
    k = {f: float(providers.simple({'slope': -1.0}).tau(F[f].pivwv() / (1. + Z))) for f in F['filters']}
    procs = max(1, min(os.cpu_count() or 1, 32))
    per_step = max(4, procs)      # galaxies per step: bounded sample of the 256-galaxy step, one per worker
    batch = synthetic.galaxy_batch(per_step, fixed_n=N_PER_GALAXY, seed=100)
    _REF.update(batch=batch, tabs=tabs, filters=F['filters'], grid={'log10age': grid['log10age'], 'log10Z': grid['log10Z']}, k=k)
    pool = mp.get_context('fork').Pool(procs) if procs > 1 else None

    def step():
        if pool is None:
            return [_ref_one_galaxy(g) for g in range(per_step)]
        return pool.map(_ref_one_galaxy, range(per_step), chunksize=1)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = min(args.steps, 10)
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = (time.time() - t0) / steps
    if pool is not None:
        pool.close()
    value = per_step * N_FILTERS / dt
    rec = {
        'impl': 'reference', 'metric': 'galaxy SEDs x filters / s', 'value': value,
        'unit': 'galaxy_SEDs_x_filters/s', 'n_gpus': world, 'steps': steps, 'warmup': max(1, min(args.warmup, 2)),
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': 'config[1] SED_fluxes: galaxies of 1e5 particles, BPASS-shape grid 51x13x%d, ISM+BC '
                               'LOS dust, %d filters; each step = %d galaxies (bounded sample of the GPU arm\'s '
                               '%d-galaxy step)' % (N_LAM, N_FILTERS, per_step, GALAXIES_PER_GPU),
                   'table_setup_s': t_setup},
        'cpu_baseline': {'value': value, 'unit': 'galaxy_SEDs_x_filters/s', 'cores': procs, 'kind': 'port',
                         'sample': 'oracle/sed_oracle.broadband (vectorised NumPy restatement of generate_Fnu), one '
                                   'galaxy per task on a pool of %d worker processes, %d galaxies per step; host has '
                                   '%d cores' % (procs, per_step, os.cpu_count())},
        'e2e': {'value': value, 'unit': 'galaxy_SEDs_x_filters/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(rec))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-imaging', action='store_true')
    ap.add_argument('--no-spectra', action='store_true')
    ap.add_argument('--no-sed', action='store_true', help='profiling aid: skip the SED timed loops')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-sed-e2e', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
