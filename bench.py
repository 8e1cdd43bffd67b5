#!/usr/bin/env python
"""bench.py -- the two SynthObs hot paths on B200, every BASELINE.json configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sections a,b,...]

Prints ONE JSON line (rank 0).  Primary record = BASELINE.json config[1] (SED_fluxes-style: galaxies of 1e5
particles, BPASS-shape grid 51x13x9000, ISM+BC LOS dust, 10 broadband filters), batched so one step streams more
than L2 from HBM.  `value` = galaxy SEDs x filters / s with inputs resident in HBM; `e2e` = the same through the
public API with pinned HOST inputs (H2D + D2H inside the timed region); `roofline` = the streaming kernel's own
launch duration (CUDA events around the launch, on its stream) against the measured HBM bandwidth.

Nested records (same line): `peaks` (TF32 / FP32-FMA / MUFU measured on this GPU), `latency_c2` (the single-galaxy
reference call), `spectra` (generate_SED + the tcgen05 contractions), `imaging` (config[2]: 1e6 particles, 8 filters,
512^2 + PSF; per-kernel rooflines, e2e through images.particle, SED -> image pipeline), `config1` (1e4 particles ->
100^2), `config4` (ONE ragged catalogue of 1e4 galaxies through the galaxy-sharded entry points, STRONG scaling,
the all_gather inside the timed region), `config5` (1e8 particles -> 8192^2 through dist.image_huge, STRONG scaling,
the NCCL reduce inside the timed region).

--impl reference times the CPU restatement of the reference (oracle/, NumPy: the reference is pure Python, there is
nothing to compile into oracle/_ref) on the same config[1] step on all host cores.
"""

import argparse
import contextlib
import io
import json
import os
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
C4_GALAXIES = 10000
C5_PARTICLES = 100000000
C5_NDIM = 8192
ALL_SECTIONS = ('peaks', 'sed', 'latency', 'spectra', 'imaging', 'c1', 'c4', 'c5', 'los')
SED_WORKLOAD = ('config[1] SED_fluxes: %d galaxies per step (and per GPU) x 1e5 particles, BPASS-shape grid 51x13x%d, '
                'ISM+BC LOS dust, %d filters' % (GALAXIES_PER_GPU, N_LAM, N_FILTERS))


def peaks_file():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            p = json.load(fh)
        return {'hbm_gbs': float(p['hbm_gbs']), 'bf16_tflops': float(p.get('bf16_tflops', 0.0)),
                'source': 'measured (MEASURED_PEAKS.json)'}
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, 5 ms period;
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


class Ctx:
    """what every section needs: rank / world, barrier, max over ranks, the measured peaks"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.args = args
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            # NCCL prints its version banner on stdout when the communicator is created: point fd 1 at stderr until
            # the first collective has run, so that stdout carries the one JSON line only
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group('nccl', device_id=self.dev)
                t = torch.zeros(1, device=self.dev)
                dist.all_reduce(t)
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
        self.peaks = peaks_file()
        self.measured = {}

    def barrier(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return float(ms)
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())


def timed_loop(ctx, step, steps):
    """K back-to-back steps between two CUDA events, bracketed by barrier + synchronize; ms per step, max over ranks"""
    import torch
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    ctx.barrier()
    return ctx.max_over_ranks(e0.elapsed_time(e1) / steps)


def timed_steps(ctx, step, steps, flush=None):
    """Mean device time of `steps` runs of `step`, each bracketed by its own CUDA events; when `flush` (a buffer
    larger than L2) is given it is rewritten between the runs, outside the timed intervals."""
    import torch
    ctx.barrier()
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
    ctx.barrier()
    return ctx.max_over_ranks(tot / steps)


def wall_loop(ctx, step, steps):
    """host wall clock per step (end-to-end calls synchronise themselves), max over ranks"""
    ctx.barrier()
    t0 = time.time()
    for _ in range(steps):
        step()
    ctx.barrier()
    return ctx.max_over_ranks((time.time() - t0) * 1e3 / steps)


def kernel_profile(step, reps=3):
    """{label: (launches per step, ms per launch)} from CUDA events recorded around every labelled launch inside the
    library (so_prof_*), on the stream the kernels run on; a separate pass, never inside a timed loop"""
    import torch
    from synthobs_b200 import _cabi
    step()
    torch.cuda.synchronize()
    _cabi.prof_enable(True)
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    rep = _cabi.prof_report()
    _cabi.prof_enable(False)
    return {k: {'launches_per_step': c / reps, 'ms_per_launch': ms / c, 'ms_per_step': ms / reps} for k, (c, ms) in rep.items()}


def setup_model(n_lam=N_LAM, n_filters=N_FILTERS, names=None):
    from synthobs_b200 import synthetic
    from synthobs_b200.sed import models
    grid = synthetic.ssp_grid(n_lam=n_lam, seed=0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = models.define_model(grid)
    model.dust_ISM = ('simple', {'slope': -1.0})
    model.dust_BC = ('simple', {'slope': -1.0})
    F = synthetic.filter_set(model.lam * (1. + Z), n_filters=n_filters, kind='gauss', lo=12000., hi=50000.,
                             prefix='JWST.FAKE.')
    if names is not None:       # rename the filters (the imaging pipeline uses the imaging filter names)
        F = {'filters': list(names), **{new: F[old] for new, old in zip(names, F['filters'])}}
    cosmo = synthetic.FixedCosmology()
    model.create_Fnu_grid(F, Z, cosmo)
    return model, F, cosmo


# ------------------------------------------------------------------ measured peaks

def run_peaks(ctx):
    """TF32 dense (cuBLAS, 8192^3), FP32 FFMA / FFMA2, MUFU.EX2 and the dust kernel's own arithmetic, measured on this
    GPU with CUDA events (same method as MEASURED_PEAKS.json: best of 5 after warm-up)."""
    import torch
    from ctypes import byref, c_double
    from synthobs_b200 import _cabi
    dev = ctx.dev
    out = torch.zeros(4, dtype=torch.float32, device=dev)

    def best_of(fn, n=5):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    res = {}
    for kind, key, scale in ((0, 'fp32_ffma_tflops', 2e-12), (1, 'fp32_ffma2_tflops', 2e-12), (2, 'mufu_ex2_per_s', 1.0),
                             (3, 'dust_arith_mufu_exp_per_s', 1.0), (4, 'dust_arith_hybrid_exp_per_s', 1.0)):
        ops = c_double(0.0)

        def fn():
            _cabi.check(_cabi.lib.so_peak_kernel(kind, 4096, 8, _cabi.ptr(out), byref(ops), _cabi.stream()), 'so_peak_kernel')
        ms = best_of(fn)
        res[key] = ops.value / (ms * 1e-3) * scale
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.rand((8192, 8192), device=dev)
    b = torch.rand((8192, 8192), device=dev)
    ms = best_of(lambda: torch.matmul(a, b))
    torch.backends.cuda.matmul.allow_tf32 = old
    res['tf32_dense_tflops'] = 2.0 * 8192 ** 3 / (ms * 1e-3) / 1e12
    del a, b
    res['hbm_gbs'] = ctx.peaks['hbm_gbs']
    res['how'] = ('so_peak_kernel: 8 CTAs x 256 threads per SM, 16 independent register chains per thread, best of 5 '
                  '(CUDA events); tf32_dense = cuBLAS torch.matmul 8192^3 with allow_tf32, best of 5; hbm_gbs from '
                  + ctx.peaks['source'])
    ctx.measured = res
    return res


# ------------------------------------------------------------------ SED (primary record)

def run_sed(ctx):
    import torch
    from synthobs_b200 import synthetic
    from synthobs_b200.sed import models
    args, dev = ctx.args, ctx.dev
    model, F, cosmo = setup_model()
    G = GALAXIES_PER_GPU
    batch = synthetic.galaxy_batch(G, fixed_n=N_PER_GALAXY, seed=100 + ctx.rank)
    keys = ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC')
    host = {k: torch.from_numpy(batch[k]).pin_memory() for k in keys}
    offsets = torch.from_numpy(batch['offsets'])
    d = {k: host[k].to(dev) for k in keys}
    d_off = offsets.to(dev)
    n_part = d['Masses'].numel()
    kw = dict(fesc=0.0, log10t_BC=7.0)

    def step_device():
        return models.generate_Fnu_batch(model, F, d['Masses'], d['Ages'], d['Metallicities'], d_off,
                                         tauVs_ISM=d['tauVs_ISM'], tauVs_BC=d['tauVs_BC'], **kw)

    def step_e2e():
        # public API with HOST (pinned) inputs: H2D of the five particle arrays + kernels + D2H of [G, F]
        return models.generate_Fnu_batch(model, F, host['Masses'], host['Ages'], host['Metallicities'], offsets,
                                         tauVs_ISM=host['tauVs_ISM'], tauVs_BC=host['tauVs_BC'], **kw)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    t_wall0 = time.time()
    ms_step = timed_loop(ctx, step_device, args.steps)
    t_wall1 = time.time()
    value = ctx.world * G * N_FILTERS / (ms_step * 1e-3)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        step_e2e()
    e2e_ms = wall_loop(ctx, step_e2e, e2e_steps)
    prof = kernel_profile(step_device)
    kname = 'sed_broadband_kernel'
    kms = prof.get(kname, {}).get('ms_per_launch', ms_step)
    algo_bytes = 40.0 * n_part                     # 5 float64 arrays read once
    achieved = algo_bytes / (kms * 1e-3) / 1e9
    hbm = ctx.peaks['hbm_gbs']
    rec = {
        'metric': 'galaxy SEDs x filters / s', 'value': value, 'unit': 'galaxy_SEDs_x_filters/s',
        'n_gpus': ctx.world, 'steps': args.steps, 'warmup': warm, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 values, f64 index compares',
        'data': 'synthetic',
        'config': {'workload': SED_WORKLOAD, 'entry_point': 'models.generate_Fnu_batch (one segmented so_sed_broadband pass)',
                   'particles_per_step_per_gpu': n_part, 'l2_policy': 'inputs (1.02 GB/step) larger than L2',
                   'parallelism': 'galaxies sharded by rank, no collective'},
        'particles_per_s': ctx.world * n_part / (ms_step * 1e-3),
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm, 'unit': 'GB/s', 'frac': achieved / hbm,
                     'traffic': 1.0369e9 if n_part == 25600000 else None,
                     'traffic_source': 'ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch '
                                       '(profiles/r02_ncu_kernels.md): 1.030 GB read + 6.9 MB written vs 1.024 GB algorithmic',
                     'peak_source': ctx.peaks['source'], 'kernel': kname, 'kernel_ms_per_launch': kms,
                     'algorithmic_bytes_per_particle': 40, 'algorithmic_bytes_per_launch': algo_bytes,
                     'step_frac': algo_bytes / (ms_step * 1e-3) / 1e9 / hbm,
                     'note': 'achieved = 40 B x particles of one launch / the broadband kernel\'s own launch duration '
                             '(CUDA events around the launch on its stream, profiling pass: the events break the '
                             'programmatic-dependent-launch overlap with the item scan and the per-galaxy reduce); '
                             'step_frac = the same bytes / the whole 3-launch step as timed for `value`'},
        'e2e': {'value': ctx.world * G * N_FILTERS / (e2e_ms * 1e-3), 'unit': 'galaxy_SEDs_x_filters/s',
                'h2d_bytes_per_step': int(sum(host[k].numel() * 8 for k in keys) + offsets.numel() * 8),
                'd2h_bytes_per_step': int(G * N_FILTERS * 8), 'ms_per_step': e2e_ms,
                'note': 'PCIe-bound: 1.02 GB per step over one Gen5 x16 link; see imaging.pipeline_e2e for the call '
                        'chain that uploads a catalogue once and keeps the per-particle fluxes on the GPU'},
        'gpu_launches': 3 * args.steps,
        'kernels': prof,
    }
    return rec, (t_wall0, t_wall1), (batch, model, F)


def run_latency(ctx):
    """config[1] exactly as examples/SED_fluxes.py:45 makes the call: ONE galaxy of 1e5 particles, NumPy arrays in,
    {filter: float} out (launch- and host-latency bound: 4 MB of inputs)."""
    from synthobs_b200 import synthetic
    from synthobs_b200.sed import models
    model, F, cosmo = setup_model()
    p = synthetic.particles(N_PER_GALAXY, seed=41)
    a = (p['Masses'], p['Ages'], p['Metallicities'])
    kw = dict(tauVs_ISM=p['tauVs_ISM'], tauVs_BC=p['tauVs_BC'], fesc=0.0)
    for _ in range(5):
        models.generate_Fnu(model, F, *a, **kw)
    n = 50
    t0 = time.time()
    for _ in range(n):
        models.generate_Fnu(model, F, *a, **kw)
    ms = (time.time() - t0) * 1e3 / n
    t0 = time.time()
    for _ in range(n):
        models.generate_Fnu_array(model, F, F['filters'][0], *a, **kw)
    ms_arr = (time.time() - t0) * 1e3 / n
    return {'call': 'models.generate_Fnu(model, F, Masses, Ages, Metallicities, tauVs_ISM=, tauVs_BC=) with NumPy inputs, '
                    '1e5 particles, 10 filters', 'ms_per_call': ms, 'galaxy_SEDs_x_filters_per_s': N_FILTERS / (ms * 1e-3),
            'generate_Fnu_array_ms_per_call': ms_arr,
            'h2d_bytes_per_call': 5 * 8 * N_PER_GALAXY, 'note': 'pageable NumPy inputs: the 4 MB upload and the host-side '
            'call overhead are the step; reference: ~10 us per particle and filter in a Python loop (SURVEY 8a)'}


# ------------------------------------------------------------------ spectra + GEMM

def run_spectra(ctx, model, F):
    """generate_SED for a batch of galaxies (five spectra over 9000 wavelengths each: segmented histogram -> 3xTF32
    tcgen05 GEMM against the SSP grid, per-particle dust kernel) followed by the second contraction spectra x filter
    weights; plus the histogram x grid GEMM alone at FLARES batch size (1e4 galaxies -> M = 2e4 rows)."""
    import torch
    from synthobs_b200 import _cabi, synthetic
    from synthobs_b200.sed import models
    args, dev = ctx.args, ctx.dev
    G = 32
    batch = synthetic.galaxy_batch(G, fixed_n=N_PER_GALAXY, seed=300 + ctx.rank)
    keys = ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC')
    d = {k: torch.from_numpy(batch[k]).to(dev) for k in keys}
    off = torch.from_numpy(batch['offsets']).to(dev)
    cosmo = synthetic.FixedCosmology()

    def step():
        o = models.sed_spectra_device(model, d['Masses'], d['Ages'], d['Metallicities'], d['tauVs_ISM'], d['tauVs_BC'],
                                      0.0, 7.0, offsets=off)
        return models.spectra_Fnu(model, F, o['total'], Z, cosmo)      # [G, F] band fluxes from the spectra

    for _ in range(3):
        step()
    steps = max(3, min(args.steps, 5))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # 256 MB > 126 MB of L2
    ms = timed_steps(ctx, step, steps, flush)
    prof = kernel_profile(step, reps=2)
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
    gms = timed_loop(ctx, gemm_only, 10)
    flops = 2.0 * M * n_lam * K
    tf32_peak = ctx.measured.get('tf32_dense_tflops') or ctx.peaks['bf16_tflops'] / 2
    src = 'measured: cuBLAS TF32 8192^3 in this run (peaks.tf32_dense_tflops)' if ctx.measured.get('tf32_dense_tflops') \
        else 'bf16 peak / 2 (peaks section skipped)'
    young = float((batch['Ages'] < 10.0).mean())
    exps = n_part * n_lam * (1.0 + young)
    dms = prof.get('dust_spectra_kernel', {}).get('ms_per_step', ms)
    exp_peak = ctx.measured.get('dust_arith_hybrid_exp_per_s') or 148 * 16 * 1.965e9
    return {
        'metric': 'galaxy SEDs/sec (5 spectra x %d wavelengths, ISM+BC dust) + %d band fluxes' % (n_lam, len(F['filters'])),
        'value': ctx.world * G / (ms * 1e-3), 'unit': 'galaxy_SEDs/s', 'ms_per_step': ms,
        'galaxy_SEDs_x_filters_per_s': ctx.world * G * len(F['filters']) / (ms * 1e-3),
        'particle_wavelengths_per_s': ctx.world * n_part * n_lam / (ms * 1e-3),
        'config': {'workload': '%d galaxies x 1e5 particles per GPU, generate_SED_batch + spectra x filters' % G,
                   'l2_policy': 'L2 flushed (256 MB write) between steps, outside the timed intervals'},
        'kernels': prof,
        'dust_roofline': {'bound': 'mufu_ex2 + fma polynomial', 'kernel': 'dust_spectra_kernel<hybrid>',
                          'achieved': exps / (dms * 1e-3), 'peak': exp_peak, 'unit': 'exp/s', 'frac': exps / (dms * 1e-3) / exp_peak,
                          'kernel_ms_per_step': dms, 'young_fraction': young,
                          'mufu_only_peak': ctx.measured.get('dust_arith_mufu_exp_per_s'),
                          'raw_mufu_ex2_per_s': ctx.measured.get('mufu_ex2_per_s'),
                          'peak_source': 'measured in this run: the kernel\'s own per-particle-wavelength arithmetic (3 MUFU '
                                         'exponentials + 1 FMA-pipe polynomial per 4) on register operands, so_peak_kernel(4)'
                                         if ctx.measured else 'nominal 148 x 16 ex2/clk x 1.965 GHz (peaks section skipped)',
                          'note': 'algorithmic exponentials: 1 per (old particle, wavelength), 2 per young one '
                                  '(models.py:285-304); duration = the dust kernel\'s own launches (CUDA events)'},
        'gemm_hist_x_grid': {'M': M, 'N': n_lam, 'K': K, 'ms': gms,
                             'l2_policy': 'operands and result (1.03 GB per launch) larger than L2',
                             'algorithmic_tflops': flops / gms / 1e9, 'executed_tf32_tflops': 3 * flops / gms / 1e9,
                             'roofline': {'bound': 'tensor', 'achieved': flops / gms / 1e9, 'peak': tf32_peak,
                                          'unit': 'TFLOP/s', 'frac': flops / gms / 1e9 / tf32_peak,
                                          'executed_frac': 3 * flops / gms / 1e9 / tf32_peak,
                                          'traffic': 1.95e9 if (M, n_lam, K) == (20000, 9000, 1326) else None,
                                          'algorithmic_bytes': 4.0 * (M * K + n_lam * K + M * n_lam),
                                          'traffic_note': 'ncu dram read 1.25 GB + write 0.70 GB per launch '
                                                          '(profiles/r02_ncu_kernels.md)',
                                          'peak_source': src,
                                          'note': 'achieved = ALGORITHMIC flops 2*M*N*K (SURVEY 8d: the x3 of 3xTF32 is '
                                                  'implementation overhead); executed_frac counts the three TF32 MMAs '
                                                  'per product'}},
    }


# ------------------------------------------------------------------ imaging (config[2])

def imaging_inputs(ctx, seed_shift=0):
    import torch
    from synthobs_b200 import synthetic
    from synthobs_b200.morph import images
    p = synthetic.particles(IMG_PARTICLES, seed=200 + ctx.rank + seed_shift, scale_kpc=3.0)
    filters = ['JWST.NIRCAM.B%d' % i for i in range(IMG_FILTERS)]
    for f in filters:
        images.filter_info[f] = {'pixel_scale': 0.031}
    cosmo = synthetic.FixedCosmology()
    width_arcsec = IMG_NDIM * 0.031 + 1e-9
    rng = np.random.default_rng(7 + ctx.rank)
    Lh = np.stack([p['Masses'] * (0.5 + rng.random(IMG_PARTICLES)) for _ in filters])
    PSFs = {f: synthetic.GaussianPSF(1.5 + 0.2 * i) for i, f in enumerate(filters)}
    return p, filters, cosmo, width_arcsec, Lh, PSFs


def imaging_stage_times(obs, Xd, Yd, Ld):
    """Device time of the stages of one imaging step (CUDA events between the stage calls; the stages are
    the same C-ABI calls particle_device_multi makes)."""
    import torch
    from synthobs_b200.morph import images
    o0 = obs[0]
    spec = images._group_spectrum(obs)

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


def count_launches(step):
    """kernel launches of one step, all libraries (torch profiler / CUPTI); None when the profiler is unavailable"""
    import torch
    try:
        from torch.profiler import ProfilerActivity, profile
        step()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        return int(sum(1 for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA
                       and 'memcpy' not in e.name.lower() and 'memset' not in e.name.lower()))
    except Exception:   # noqa: BLE001
        return None


def run_imaging(ctx):
    import torch
    from synthobs_b200.morph import images
    from synthobs_b200.sed import models
    args, dev = ctx.args, ctx.dev
    p, filters, cosmo, width_arcsec, Lh, PSFs = imaging_inputs(ctx)
    Xh = torch.from_numpy(p['X']).pin_memory()
    Yh = torch.from_numpy(p['Y']).pin_memory()
    Lh_t = torch.from_numpy(Lh).pin_memory()
    Xd, Yd, Ld = Xh.to(dev), Yh.to(dev), Lh_t.to(dev)
    kw = dict(smoothing=('adaptive', float(IMG_K)), super_sampling=IMG_SS)
    obs = [images.observed(f, cosmo, Z, width_arcsec, PSF=PSFs[f], **kw) for f in filters]
    assert obs[0].ndim == IMG_NDIM, obs[0].ndim
    stats = {}

    def step_device():
        mx, my = images.median_device_pair(Xd, Yd)    # recentring (images.py:183-184, np.median semantics)
        r = images.particle_device_multi(obs, Xd - mx, Yd - my, Ld)
        stats.update(r['stats'])
        return r['data']

    Ldict = {f: Lh_t[i] for i, f in enumerate(filters)}
    Xw, Yw = Xh.clone().pin_memory(), Yh.clone().pin_memory()

    def step_e2e():
        # THE reference-named call (images.py:263): pinned host X, Y, {filter: L} in, {filter: img} (NumPy) out.  The
        # call recentres X, Y in place (as upstream does), so from the second call on the catalogue it is handed is
        # already centred: same work, same images
        return images.particle(Xw, Yw, Ldict, filters, cosmo, Z, width_arcsec, PSFs=PSFs, **kw)

    for _ in range(3):
        step_device()
    steps = max(3, min(args.steps, 10))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # inputs (80 MB) fit in L2: flush between steps
    ms = timed_steps(ctx, step_device, steps, flush)
    del flush
    for _ in range(2):
        step_e2e()
    e2e_ms = wall_loop(ctx, step_e2e, steps)
    stages = imaging_stage_times(obs, Xd, Yd, Ld)
    prof = kernel_profile(step_device)
    # what the deposit's multiply-add loop really evaluates (so_img_deposit_count: a counting instantiation of the kernel)
    from synthobs_b200 import _cabi
    dep_cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    _cabi.check(_cabi.lib.so_img_deposit_count(_cabi.ptr(dep_cnt)), 'so_img_deposit_count')
    try:
        step_device()
        torch.cuda.synchronize()
    finally:
        _cabi.lib.so_img_deposit_count(None)
    dep_visits, dep_slots = (float(v) for v in dep_cnt.tolist())

    # SED -> images pipeline (examples/morph_image_observed.py:93-95): the catalogue goes up ONCE (56 B/particle), the
    # per-particle fluxes of the 8 filters never leave the GPU, 8 images come back
    pmodel, pF, _ = setup_model(n_filters=IMG_FILTERS, names=filters)
    cat = {k: torch.from_numpy(p[k]).pin_memory() for k in ('X', 'Y', 'Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC')}

    def step_pipeline():
        c = {k: v.to(dev, non_blocking=True) for k, v in cat.items()}
        Lp = models.generate_Fnu_arrays(pmodel, pF, c['Masses'], c['Ages'], c['Metallicities'], tauVs_ISM=c['tauVs_ISM'],
                                        tauVs_BC=c['tauVs_BC'])
        out = images.particle(c['X'], c['Y'], Lp, filters, cosmo, Z, width_arcsec, PSFs=PSFs, **kw)
        return torch.stack([out[f].data for f in filters]).cpu()
    for _ in range(2):
        step_pipeline()
    pipe_ms = wall_loop(ctx, step_pipeline, steps)
    launches = count_launches(step_device)

    deposits = stats.get('deposits', 0.0)
    pairs = stats.get('pairs', 0.0)
    nsel = stats.get('selected', 0.0)
    ref_equiv = nsel * (IMG_NDIM * IMG_SS) ** 2 * IMG_FILTERS
    hbm = ctx.peaks['hbm_gbs']
    ffma = ctx.measured.get('fp32_ffma2_tflops') or ctx.measured.get('fp32_ffma_tflops')

    def kms(name):
        return prof.get(name, {}).get('ms_per_step')
    kern = {}
    dep_ms = kms('deposit_kernel')
    if dep_ms:
        useful = 2.0 * deposits * IMG_FILTERS                      # 2 flop per in-support (particle, pixel, channel) deposit
        # a (lane, particle) visit of the multiply-add loop evaluates one 2 x 4 pixel block for all channels; a lane only
        # visits the staged particles whose support box meets its block (round 1 evaluated 256 pixels per pair)
        executed = 2.0 * dep_visits * 8 * IMG_FILTERS
        slots = 2.0 * dep_slots * 8 * IMG_FILTERS                  # issue slots: every round occupies all 32 lanes
        kern['deposit_kernel'] = {
            'bound': 'fp32_fma', 'ms_per_step': dep_ms, 'algorithmic_flops': useful, 'executed_flops': executed,
            'achieved': useful / (dep_ms * 1e-3) / 1e12, 'executed_tflops': executed / (dep_ms * 1e-3) / 1e12,
            'peak': ffma, 'unit': 'TFLOP/s', 'frac': (useful / (dep_ms * 1e-3) / 1e12 / ffma) if ffma else None,
            'executed_frac': (executed / (dep_ms * 1e-3) / 1e12 / ffma) if ffma else None,
            'useful_fma_fraction': useful / executed if executed else None,
            'issue_slot_flops': slots, 'lanes_busy_fraction': executed / slots if slots else None,
            'round1_executed_flops': 2.0 * pairs * 256 * IMG_FILTERS,
            'algorithmic_bytes': IMG_PARTICLES * (8 + 8 + 4 + 4 * IMG_FILTERS) + 4 * IMG_FILTERS * (IMG_NDIM * IMG_SS) ** 2,
            'hbm_frac': (IMG_PARTICLES * (20 + 4 * IMG_FILTERS) + 4 * IMG_FILTERS * (IMG_NDIM * IMG_SS) ** 2) / (dep_ms * 1e-3) / 1e9 / hbm,
            'peak_source': 'measured FFMA2 throughput (peaks.fp32_ffma2_tflops)' if ffma else None,
            'note': 'SURVEY 8d: binding roof = max(bytes/HBM, flops/FP32-FMA, exps/MUFU); at ~18 pixels per particle '
                    'the algorithmic work is tiny against either roof; executed_flops = multiply-adds of the (lane, particle) '
                    'visits counted by the kernel itself (so_img_deposit_count), issue_slot_flops = the same for rounds x 32 lanes'}
    for name in ('knn_query_kernel', 'knn_group_kernel'):
        q_ms = kms(name)
        if q_ms:
            kern[name] = {'bound': 'hbm (algorithmic) / latency (actual)', 'ms_per_step': q_ms,
                          'algorithmic_bytes': 20.0 * nsel, 'achieved': 20.0 * nsel / (q_ms * 1e-3) / 1e9, 'peak': hbm,
                          'unit': 'GB/s', 'frac': 20.0 * nsel / (q_ms * 1e-3) / 1e9 / hbm,
                          'queries_per_s': nsel / (q_ms * 1e-3),
                          'note': 'SURVEY 8d: 16 B in + 4 B out per particle; a tree search is latency- and '
                                  'instruction-bound, the HBM fraction is reported because 8d asks for it'}
    algo_bytes = IMG_PARTICLES * (16 + 8 * IMG_FILTERS) + 4 * IMG_FILTERS * ((IMG_NDIM * IMG_SS) ** 2 + IMG_NDIM ** 2)
    top = max(kern.items(), key=lambda kv: kv[1]['ms_per_step'])[0] if kern else None
    return {
        'metric': 'particles/sec imaged', 'unit': 'particles/s (all %d filter images)' % IMG_FILTERS,
        'value': ctx.world * IMG_PARTICLES / (ms * 1e-3), 'ms_per_step': ms,
        'particle_filter_images_per_s': ctx.world * IMG_PARTICLES * IMG_FILTERS / (ms * 1e-3),
        'deposits_evaluated_per_s': ctx.world * deposits * IMG_FILTERS / (ms * 1e-3),
        'reference_equivalent_deposits_per_s': ctx.world * ref_equiv / (ms * 1e-3),
        'roofline': dict(kern[top], kernel=top) if top else None,
        'kernel_rooflines': kern,
        'whole_step_hbm_frac': algo_bytes / (ms * 1e-3) / 1e9 / hbm,
        'stages_ms': stages, 'kernels': prof, 'launches_per_step': launches,
        'config': {'workload': 'config[2] observed-frame: %d particles, %d filters, %dx%d, super_sampling %d, '
                               "('adaptive', %d), Gaussian PSF FFT convolution, rebin" %
                               (IMG_PARTICLES, IMG_FILTERS, IMG_NDIM, IMG_NDIM, IMG_SS, IMG_K),
                   'l2_policy': 'L2 flushed (256 MB write) between steps, outside the timed intervals',
                   'support_sigmas': images.SUPPORT_SIGMAS, 'pairs_particle_tile': pairs,
                   'selected': nsel, 'deposits_per_filter': deposits},
        'e2e': {'value': ctx.world * IMG_PARTICLES / (e2e_ms * 1e-3), 'unit': 'particles/s', 'ms_per_step': e2e_ms,
                'entry_point': 'images.particle(X, Y, {filter: L}, filters, cosmo, z, width, smoothing=, PSFs=, super_sampling=) '
                               'with pinned CPU tensors; returns {filter: img} NumPy; X, Y recentred in place',
                'h2d_bytes_per_step': int(IMG_PARTICLES * 8 * (2 + IMG_FILTERS)),
                'd2h_bytes_per_step': int(2 * IMG_FILTERS * IMG_NDIM * IMG_NDIM * 8 + 16 * IMG_PARTICLES)},
        'pipeline_e2e': {'value': ctx.world * IMG_PARTICLES / (pipe_ms * 1e-3), 'unit': 'particles/s', 'ms_per_step': pipe_ms,
                         'chain': 'pinned catalogue (X, Y, Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC) -> H2D once -> '
                                  'models.generate_Fnu_arrays (8 filters, stays on the GPU) -> images.particle -> 8 images D2H',
                         'h2d_bytes_per_step': int(IMG_PARTICLES * 56),
                         'd2h_bytes_per_step': int(IMG_FILTERS * IMG_NDIM * IMG_NDIM * 4)},
    }


def run_los(ctx):
    """Line-of-sight metal column density (synthobs_b200.los, so_los_column): one object of 1e5 star and 1e5
    gas particles per GPU = 1e10 (star, gas) pairs per step.  Compute-bound pair kernel: reported against the measured
    FP32 FMA roof with 7 flops per pair (2 subtractions, 2 multiply-adds, 1 multiply: what every pair costs before the
    support test); the oracle (NumPy, float64) timed on a sample of the same stars is the CPU baseline."""
    import torch
    from synthobs_b200 import los
    dev = ctx.dev
    ns = ng = 100000
    g = torch.Generator(device=dev)
    g.manual_seed(7 + ctx.rank)
    stars = [1.5 * torch.randn(ns, dtype=torch.float64, device=dev, generator=g) for _ in range(3)]
    gas = [2.0 * torch.randn(ng, dtype=torch.float64, device=dev, generator=g) for _ in range(3)]
    m = 1e6 * (0.5 + torch.rand(ng, dtype=torch.float64, device=dev, generator=g))
    Z = 10 ** (-4 + 2.3 * torch.rand(ng, dtype=torch.float64, device=dev, generator=g))
    h = 10 ** (-1.3 + 1.3 * torch.rand(ng, dtype=torch.float64, device=dev, generator=g))

    def step():
        return los.metal_surface_density(*stars, *gas, m, Z, h, scale=1e-6)                # default: the cell grid at this size

    def step_all():
        return los.metal_surface_density(*stars, *gas, m, Z, h, scale=1e-6, grid=False)    # every star against all the gas
    for _ in range(3):
        out = step()
        step_all()
    ms = timed_loop(ctx, step, max(3, min(ctx.args.steps, 10)))
    ms_all = timed_loop(ctx, step_all, max(3, min(ctx.args.steps, 10)))
    plan = los._cell_grid(stars[0], stars[1], gas[0], gas[1], h)
    tested = float(np.sum(np.diff(plan[1]) * np.diff(plan[3]))) if plan is not None else float(ns) * ng
    pairs = float(ns) * ng
    ffma = ctx.measured.get('fp32_ffma_tflops')
    rec = {'metric': 'star-gas pairs / s (line-of-sight metal column density)', 'unit': 'pairs/s',
           'value': ctx.world * pairs / (ms * 1e-3), 'ms_per_step': ms, 'stars_per_s': ctx.world * ns / (ms * 1e-3),
           'pairs_tested_with_the_cell_grid': tested, 'cell_grid_cells_with_stars': int((np.diff(plan[1]) > 0).sum()) if plan is not None else None,
           'all_pairs': {'ms_per_step': ms_all, 'value': ctx.world * pairs / (ms_all * 1e-3)},
           'config': {'workload': '1 object per GPU: %d star x %d gas particles, smoothing lengths 0.05..1 (extent ~2); value = '
                                  'pairs of the specification (N_star x N_gas) per second, whatever subset the grid tests' % (ns, ng),
                      'l2_policy': 'inputs 6.4 MB: L2-resident by design (every star reads its gas list)'},
           'roofline': {'bound': 'fp32_fma', 'algorithmic_flops': 7.0 * pairs, 'achieved': 7.0 * pairs / (ms_all * 1e-3) / 1e12,
                        'peak': ffma, 'unit': 'TFLOP/s', 'frac': (7.0 * pairs / (ms_all * 1e-3) / 1e12 / ffma) if ffma else None,
                        'kernel': 'los_kernel (all pairs, grid=False)',
                        'peak_source': 'measured FFMA throughput (peaks.fp32_ffma_tflops)' if ffma else None},
           'mean_column_Msol_pc2': float(out.mean())}
    if ctx.rank == 0 and ctx.world == 1 and not ctx.args.no_cpu:
        from oracle import los_oracle
        h_stars = [a[:512].cpu().numpy() for a in stars]
        h_gas = [a.cpu().numpy() for a in gas] + [m.cpu().numpy(), Z.cpu().numpy(), h.cpu().numpy()]
        t0 = time.time()
        ref = los_oracle.metal_surface_density(*h_stars, *h_gas, scale=1e-6)
        dt = time.time() - t0
        rec['cpu_baseline'] = {'value': 512.0 * ng / dt, 'unit': 'pairs/s', 'cores': 1, 'kind': 'port',
                               'sample': 'oracle/los_oracle.metal_surface_density (the specification itself; the reference has no '
                                         'implementation) for the first 512 stars against all %d gas particles in %.1f s' % (ng, dt),
                               'max_abs_diff_over_max': float(np.max(np.abs(out[:512].cpu().numpy() - ref)) / np.max(ref))}
    return rec


def run_c1(ctx):
    """config[0]: examples/morph_image_physical.py-style, 1e4 particles onto a 100 x 100 physical image through
    images.core, the three smoothing modes; NumPy in / NumPy out, ms per call."""
    from synthobs_b200 import synthetic
    from synthobs_b200.morph import images
    p = synthetic.particles(10000, seed=5)
    out = {}
    for name, sm in (('ngp', False), ('convolved_gaussian', ('convolved_gaussian', 0.2)), ('adaptive16', ('adaptive', 16))):
        for _ in range(3):
            images.core(p['X'], p['Y'], p['Masses'], resolution=0.1, ndim=100, smoothing=sm)
        n = 20
        t0 = time.time()
        for _ in range(n):
            img = images.core(p['X'], p['Y'], p['Masses'], resolution=0.1, ndim=100, smoothing=sm)
        ms = (time.time() - t0) * 1e3 / n
        out[name] = {'ms_per_call': ms, 'particles_per_s': 10000 / (ms * 1e-3)}
        if name == 'adaptive16':
            out[name]['deposits_evaluated_per_s'] = img.info['deposits'] / (ms * 1e-3)
            out[name]['reference_equivalent_deposits_per_s'] = img.info['selected'] * 1e4 / (ms * 1e-3)
    rec = {'workload': 'config[0]: 1e4 particles -> 100 x 100 physical image, images.core (host arrays in and out)', **out}
    if ctx.rank == 0 and ctx.world == 1 and not ctx.args.no_cpu:
        from oracle import image_oracle
        t0 = time.time()
        image_oracle.core(p['X'], p['Y'], p['Masses'], resolution=0.1, ndim=100, smoothing=('adaptive', 16))
        rec['cpu_baseline'] = {'value': 10000 / (time.time() - t0), 'unit': 'particles/s', 'cores': 1, 'kind': 'port',
                               'sample': 'the whole config: oracle image_oracle.core adaptive 16 (cKDTree + separable NumPy deposit)'}
    return rec


# ------------------------------------------------------------------ config 4: one ragged catalogue, galaxy-sharded

def run_c4(ctx, model, F):
    """ONE catalogue of 1e4 galaxies (sizes 10^U(3,5), counter-based particles: every rank generates exactly its share)
    through dist.generate_Fnu_sharded and dist.core_batch_sharded: LPT shares, the all_gather inside the timed
    region; total work fixed as N grows (strong scaling)."""
    import torch
    from synthobs_b200 import dist as sdist, synthetic
    args, dev = ctx.args, ctx.dev
    G = args.c4_galaxies
    sizes, offsets = synthetic.ragged_sizes(G, 1000, 100000, seed=2)
    mine, local_off = sdist.shard_plan(offsets)
    ids = torch.cat([torch.arange(offsets[g], offsets[g + 1], dtype=torch.int64, device=dev) for g in mine]) \
        if len(mine) else torch.zeros(0, dtype=torch.int64, device=dev)
    cat = synthetic.hashed_catalogue(ids)
    del ids
    shards = sdist.shard_galaxies(sizes, ctx.world)
    share = {'indices': mine, 'offsets': torch.from_numpy(local_off).to(dev), 'n_total': G, 'all_indices': shards}   # resident
    share.update({k: cat[k] for k in sdist.SED_KEYS})
    n_local = int(local_off[-1])
    n_total = int(offsets[-1])

    def sed_step():
        return sdist.generate_Fnu_sharded(model, F, None, share=share)
    for _ in range(3):
        rows = sed_step()
    assert tuple(rows.shape) == (G, len(F['filters']))
    steps = max(3, min(args.steps, 10))
    sed_ms = timed_loop(ctx, sed_step, steps)
    # end to end: the share as pinned host arrays, uploaded inside the call
    hshare = {'indices': mine, 'offsets': local_off, 'n_total': G, 'all_indices': shards}
    hshare.update({k: cat[k].cpu().pin_memory() for k in sdist.SED_KEYS})

    def sed_e2e():
        return sdist.generate_Fnu_sharded(model, F, None, share=hshare).cpu()
    sed_e2e()
    sed_e2e_ms = wall_loop(ctx, sed_e2e, 3)
    del hshare
    # images: 100 x 100 physical, ('adaptive', 16), every object of the share in segmented passes
    ishare = {'indices': mine, 'offsets': local_off, 'n_total': G, 'X': cat['X'], 'Y': cat['Y'],
              'L': cat['Masses'] * cat['weight']}

    def img_step():
        return sdist.core_batch_sharded(None, None, None, None, resolution=0.1, ndim=100, smoothing=('adaptive', 16),
                                        share=ishare, gather=False)
    img_step()
    img_ms = timed_loop(ctx, img_step, 2)
    loads = ctx.max_over_ranks(n_local)
    return {'workload': 'config[3] FLARES-scale batch: ONE catalogue of %d galaxies x 1e3..1e5 particles (%d particles), LPT '
                        'shares, %d filters + 100x100 adaptive-16 images' % (G, n_total, len(F['filters'])),
            'scaling': 'strong', 'n_gpus': ctx.world, 'galaxies': G, 'particles': n_total,
            'max_share_particles': loads, 'share_imbalance': loads / (n_total / ctx.world),
            'sed': {'entry_point': 'dist.generate_Fnu_sharded (segmented so_sed_broadband on the share + all_gather of [G, F])',
                    'ms_per_step': sed_ms, 'value': G * len(F['filters']) / (sed_ms * 1e-3), 'unit': 'galaxy_SEDs_x_filters/s',
                    'particles_per_s': n_total / (sed_ms * 1e-3),
                    'hbm_frac_of_n_gpus': 40.0 * n_total / (sed_ms * 1e-3) / 1e9 / (ctx.peaks['hbm_gbs'] * ctx.world),
                    'e2e': {'ms_per_step': sed_e2e_ms, 'value': G * len(F['filters']) / (sed_e2e_ms * 1e-3),
                            'h2d_bytes_per_step_per_rank': 40 * n_local, 'd2h_bytes_per_step': G * len(F['filters']) * 8}},
            'images': {'entry_point': 'dist.core_batch_sharded(..., gather=False): images.core of every object of the share',
                       'ms_per_step': img_ms, 'value': n_total / (img_ms * 1e-3), 'unit': 'particles/s',
                       'galaxies_per_s': G / (img_ms * 1e-3)}}


# ------------------------------------------------------------------ config 5: one huge image, particle shares + NCCL

def run_c5(ctx):
    """1e8 particles onto 8192 x 8192, ('adaptive', 8), through dist.image_huge: every rank answers the k-NN and
    deposits for its Morton-range share, the partial images are reduced over NVLink inside the timed region."""
    import torch
    from synthobs_b200 import dist as sdist
    args, dev = ctx.args, ctx.dev
    N, ndim, S = args.c5_particles, C5_NDIM, 10.0
    g = torch.Generator(device=dev)
    g.manual_seed(99)                      # same stream on every rank: every rank holds the whole catalogue
    X = S * torch.randn(N, dtype=torch.float64, device=dev, generator=g)
    Y = S * torch.randn(N, dtype=torch.float64, device=dev, generator=g)
    m = N // 20
    X[:m] = 3.0 + 0.5 * torch.randn(m, dtype=torch.float64, device=dev, generator=g)      # a dense clump
    Y[:m] = -7.0 + 0.5 * torch.randn(m, dtype=torch.float64, device=dev, generator=g)
    L = 8.47e5 * (0.5 + torch.rand(N, dtype=torch.float64, device=dev, generator=g))
    res = 8.0 * S / ndim

    def step():
        return sdist.image_huge(X, Y, L, res, ndim, IMG_K)
    # warm-up: the share fractions settle on the measured busy times of the ranks and freeze, then both ways of summing
    # the partial images are timed by the library and the faster one is kept (dist.reduce_mode)
    for _ in range(sdist.SHARE_ADAPT_CALLS + len(sdist.REDUCE_TRIALS) + 2 if ctx.world > 1 else 3):
        img = step()
    flux = float(img.double().sum())
    ms = timed_loop(ctx, step, max(3, min(args.steps, 5)))
    stages = {}
    for _ in range(2):
        sdist.image_huge(X, Y, L, res, ndim, IMG_K, stages=stages)
    stages = {k: ctx.max_over_ranks(v / 2) for k, v in stages.items()}
    chosen, forced = 'none (one rank)', {}
    if ctx.world > 1:
        bal = sdist._share_balance.get((N, int(ndim), int(IMG_K), ctx.world), {})
        env0 = os.environ.get('SO_PEER_REDUCE')
        chosen = {'peer': 'peer-memory block exchange fused into the deposit (csrc/peer.cu)', 'nccl': 'ncclAllReduce of the image',
                  None: 'forced by SO_PEER_REDUCE=%s' % env0}[bal.get('mode') if sdist.reduce_mode() == 'auto' else None]
        # the same step with each way forced
        try:
            for name, v in (('peer', '1'), ('nccl', '0')):
                os.environ['SO_PEER_REDUCE'] = v
                for _ in range(2):
                    step()
                forced[name] = timed_loop(ctx, step, max(3, min(args.steps, 5)))
        finally:
            if env0 is None:
                os.environ.pop('SO_PEER_REDUCE', None)
            else:
                os.environ['SO_PEER_REDUCE'] = env0
    sel = float(((X.abs() < 4 * S) & (Y.abs() < 4 * S)).sum())
    return {'workload': 'config[4] single huge object: %d particles -> %d x %d, (\'adaptive\', %d), Morton-range particle '
                        'shares, partial images summed over NVLink' % (N, ndim, ndim, IMG_K),
            'entry_point': 'dist.image_huge', 'scaling': 'strong', 'n_gpus': ctx.world, 'ms_per_step': ms,
            'value': N / (ms * 1e-3), 'unit': 'particles/s', 'selected': sel,
            'flux_conservation': flux / float(L[(X.abs() < 4 * S) & (Y.abs() < 4 * S)].sum()) - 1.0,
            'stages_ms_max_over_ranks': stages, 'image_bytes': 4 * ndim * ndim,
            'reduce': chosen, 'ms_per_step_with_peer_exchange': forced.get('peer'),
            'ms_per_step_with_nccl_allreduce': forced.get('nccl')}


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


# ------------------------------------------------------------------ ours (GPU)

def run_ours(args):
    import torch.distributed as dist
    from synthobs_b200 import _cabi
    ctx = Ctx(args)
    _cabi.require_device()
    sections = set(args.sections.split(',')) if args.sections else set(ALL_SECTIONS)
    sampler = ClockSampler(ctx.local)
    sampler.start()
    t_start = time.time()
    rec = {}
    peaks = run_peaks(ctx) if 'peaks' in sections else None
    batch = model = F = None
    if 'sed' in sections:
        rec, _, (batch, model, F) = run_sed(ctx)
    else:
        rec = {'metric': 'galaxy SEDs x filters / s', 'value': None, 'unit': 'galaxy_SEDs_x_filters/s', 'n_gpus': ctx.world,
               'note': 'SED section skipped (--sections)'}
        model, F, _ = setup_model()
    rec['peaks'] = peaks
    if 'latency' in sections and ctx.rank == 0:
        rec['latency_c2'] = run_latency(ctx)
    ctx.barrier()
    if 'spectra' in sections:
        rec['spectra'] = run_spectra(ctx, model, F)
    if 'imaging' in sections:
        rec['imaging'] = run_imaging(ctx)
    if 'c1' in sections and ctx.rank == 0:
        rec['config1'] = run_c1(ctx)
    ctx.barrier()
    if 'c4' in sections:
        rec['config4'] = run_c4(ctx, model, F)
    if 'c5' in sections:
        rec['config5'] = run_c5(ctx)
    if 'los' in sections:
        rec['los'] = run_los(ctx)
    sampler.stop()
    rec['clocks'] = sampler.summary(t_start, time.time())     # all timed regions of this run
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu:
        if batch is not None:
            rec['cpu_baseline'] = cpu_baseline_sed(batch, model, F, budget_s=12.0)
        if rec.get('imaging'):
            rec['imaging']['cpu_baseline'] = cpu_baseline_imaging(budget_s=12.0)
    else:
        rec.setdefault('cpu_baseline', None)
    if ctx.rank == 0:
        print(json.dumps(rec))
    if ctx.world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ reference arm

_REF = {}


def _ref_task(task):
    from oracle import sed_oracle
    gi, half = task
    b, off = _REF['batch'], _REF['batch']['offsets']
    s, e = off[gi], off[gi + 1]
    mid = (s + e) // 2
    s, e = (s, mid) if half == 0 else (mid, e)
    out = sed_oracle.broadband(_REF['tabs'], _REF['filters'], _REF['grid']['log10age'], _REF['grid']['log10Z'],
                               b['Masses'][s:e], b['Ages'][s:e], b['Metallicities'][s:e], tauVs_ISM=b['tauVs_ISM'][s:e],
                               tauVs_BC=b['tauVs_BC'][s:e], k_ISM=_REF['k'], k_BC=_REF['k'], fesc=0.0)
    return [out[f] for f in _REF['filters']]


def run_reference(args):
    """The reference's own CPU path for the headline metric, on the host cores of the box: the oracle port
    (vectorised NumPy restatement of generate_Fnu; the reference is pure Python, so there is no oracle/_ref to
    compile), half a galaxy per task over a pool of worker processes on all host cores.  Same config[1] step as the
    GPU arm (GALAXIES_PER_GPU galaxies of 1e5 particles) unless --ref-galaxies bounds it; --steps / --warmup honoured."""
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
    procs = max(1, os.cpu_count() or 1)
    per_step = args.ref_galaxies
    batch = synthetic.galaxy_batch(per_step, fixed_n=N_PER_GALAXY, seed=100)
    _REF.update(batch=batch, tabs=tabs, filters=F['filters'], grid={'log10age': grid['log10age'], 'log10Z': grid['log10Z']}, k=k)
    tasks = [(g, h) for g in range(per_step) for h in (0, 1)]
    pool = mp.get_context('fork').Pool(procs) if procs > 1 else None

    def step():
        if pool is None:
            return [_ref_task(t) for t in tasks]
        return pool.map(_ref_task, tasks, chunksize=max(1, len(tasks) // (4 * procs)))

    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / args.steps
    if pool is not None:
        pool.close()
    value = per_step * N_FILTERS / dt
    sample = ('the whole %d-galaxy step' % per_step) if per_step == GALAXIES_PER_GPU else \
        ('%d of the %d galaxies of the step (throughput-normalised)' % (per_step, GALAXIES_PER_GPU))
    rec = {
        'impl': 'reference', 'metric': 'galaxy SEDs x filters / s', 'value': value,
        'unit': 'galaxy_SEDs_x_filters/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': SED_WORKLOAD, 'entry_point': 'oracle/sed_oracle.broadband (kind: port)',
                   'galaxies_per_step': per_step, 'table_setup_s': t_setup},
        'cpu_baseline': {'value': value, 'unit': 'galaxy_SEDs_x_filters/s', 'cores': procs, 'kind': 'port',
                         'sample': '%s; oracle/sed_oracle.broadband (vectorised NumPy restatement of generate_Fnu -- the '
                                   'builder\'s port, not code from the reference tree, which is a per-particle Python loop '
                                   '~100x slower), half a galaxy per task on a pool of %d worker processes; host has %d '
                                   'cores' % (sample, procs, os.cpu_count())},
        'e2e': {'value': value, 'unit': 'galaxy_SEDs_x_filters/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(rec))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--sections', default='', help='comma list of %s (default: all)' % ','.join(ALL_SECTIONS))
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--c4-galaxies', type=int, default=C4_GALAXIES)
    ap.add_argument('--c5-particles', type=int, default=C5_PARTICLES)
    ap.add_argument('--ref-galaxies', type=int, default=GALAXIES_PER_GPU,
                    help='galaxies per reference-arm step (default: the whole GPU step)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
