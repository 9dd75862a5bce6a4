"""Throughput of the per-spectrum likelihood hot path on B200 (and, with
--impl reference, of its CPU restatement on the host cores).

    python bench.py --gpus N --steps K --warmup W [--workload desi|sdss|test]

One "step" = one pass of the hot path over one batch of synthetic spectra:
per spectrum an RV-grid chi-square scan (arange(min_vel, max_vel, 5) trials of
one template, + find_best statistics) followed by the optimiser-phase
evaluations (each a NEW template: interpolate, broaden, spline, resample,
continuum solve, chi-square).  See DESIGN.md "Measurement" for what the step
contains in this round.  Prints ONE JSON line (rank 0).
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

from rvspecfit_b200 import synth  # noqa: E402

WORKLOADS = {
    # three DESI arms, 40x11x13x5 polylinear grid
    'desi': dict(arms=('desi_b', 'desi_r', 'desi_z'), layout='desi', npoly=10,
                 min_vel=-1500, max_vel=1500),
    # config 2: SDSS/BOSS single arm, 24x9x9x4 grid
    'sdss': dict(arms=('sdss',), layout='sdss', npoly=10, min_vel=-1500, max_vel=1500),
    # config 4: Gaia RVS window, FFT cross-correlation first guess over a (template x
    # vsini) bank at npoints 8192 (gaia_rvs/make_gaia.sh:5-11, make_ccf.py:496-497,556)
    'gaia_rvs': dict(arms=('gaiarvs',), layout='small', npoly=10, min_vel=-1000, max_vel=1000,
                     ccf=dict(npoints=8192, every=7, vsinis=(0., 10., 30., 100., 300.))),
    # config 1 shape (correctness-sized; fits L2, not a roofline workload)
    'test': dict(arms=('test',), layout='test', npoly=15, min_vel=-1000, max_vel=1000),
}
# evaluations of one reference vel_fit.process call on a DESI-shaped object
# (SURVEY.md section 6/8d probe): 2866 get_chisq, 1293 of them with a new template
EVALS_PER_FIT = 1293
# starting point of every fit in --mode fit (the reference starts from the CCF
# first guess; a fixed mid-grid start exercises the optimiser at least as hard)
FIT_START = {'teff': 5500., 'logg': 3.0, 'feh': -1.0, 'alpha': 0.2, 'vsini': 10.}


def make_config(w):
    return dict(min_vel=w['min_vel'], max_vel=w['max_vel'], vel_step0=5, max_vsini=500,
                min_vsini=0.1, min_vel_step=0.2, second_minimizer=True,
                template_lib='synthetic/')


def workload_string(wname):
    """config.workload of both arms (GPU and --impl reference): the same string."""
    w = WORKLOADS[wname]
    g = synth.GRIDS[w['layout']]
    nodes = int(np.prod([len(g[k]) for k in synth.PARNAMES]))
    return (f'{wname}: arms {list(w["arms"])}, grid {w["layout"]} ({nodes} nodes, fp32), '
            f'npoly {w["npoly"]}, RV grid [{w["min_vel"]}, {w["max_vel"]}) step 5 km/s')


def grid_cache_path(wname, arm):
    d = '/dev/shm' if os.path.isdir('/dev/shm') else '/tmp'
    return os.path.join(d, f'rvs_bench_grid_{wname}_{arm}.npy')


def make_inputs(wname, nspec, seed, mmap_grids=False):
    """Synthetic banks + spectra (host numpy).  Objects are lists of
    (name, lam, spec, espec, badmask).  With mmap_grids the template rows are
    read from the files the parent process wrote (shared page cache, like the
    reference's mmap of interpdat_<setup>.npy, spec_inter.py:353-355)."""
    w = WORKLOADS[wname]
    setups = []
    for k, a in enumerate(w['arms']):
        dats = np.load(grid_cache_path(wname, a), mmap_mode='r') if mmap_grids else None
        setups.append(synth.make_setup(a, w['layout'], seed=21 + k, dats=dats))
    pars = synth.random_params(w['layout'], nspec, seed)
    rs = np.random.RandomState(seed + 1)
    vel = rs.normal(0, 150., nspec)
    sn = np.exp(rs.uniform(np.log(5), np.log(100), nspec))
    arms = [synth.fast_spectra(st, pars, vel, sn, seed + 10 + k) for k, st in enumerate(setups)]
    objects = [[(st['name'], a[0], a[1][i], a[2][i], a[3][i]) for st, a in zip(setups, arms)]
               for i in range(nspec)]
    return setups, objects, pars, vel


def trial_points(pars, vel, layout, nevals, seed):
    """Optimiser-like trial points around the truth: (nevals, B, 4) params,
    (nevals, B) velocities and vsini."""
    rs = np.random.RandomState(seed)
    B = len(vel)
    g = synth.GRIDS[layout]
    lo = np.array([g[k][0] for k in synth.PARNAMES])
    hi = np.array([g[k][-1] for k in synth.PARNAMES])
    scale = np.array([150., 0.25, 0.2, 0.1])
    p = pars[None] + scale * rs.normal(size=(nevals, B, 4))
    eps = 1e-3 * (hi - lo)
    p = np.clip(p, lo + eps, hi - eps)
    v = vel[None] + 3.0 * rs.normal(size=(nevals, B))
    vs = np.abs(10 + 8 * rs.normal(size=(nevals, B)))
    return p, v, vs


def algorithmic_bytes(setups, objects, nvert=16):
    """nv * sum_arms Npix_t * sizeof(grid) + sum_arms Npix_obs*16 + 8."""
    npt = sum(len(s['lam']) for s in setups)
    npo = sum(len(a[1]) for a in objects[0])
    return nvert * npt * 4 + npo * 16 + 8, npt, npo


# ------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for t, line in self.rows:
            if not (t0 <= t <= t1 + 0.3):
                continue
            f = [x.strip() for x in line.split(',')]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, val in zip(names, f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return None
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm))


# ------------------------------------------------------------------ CCF workload
def ccf_bank_path(wname):
    d = '/dev/shm' if os.path.isdir('/dev/shm') else '/tmp'
    return os.path.join(d, f'rvs_bench_ccfbank_{wname}.npz')


def ccf_bank(wname, setup):
    """The CCF template bank of the workload (built once per box with the device routines
    of this package, kept in /dev/shm so that the CPU arm's processes map the same arrays)."""
    from rvspecfit_b200 import make_ccf, spec_inter
    w = WORKLOADS[wname]
    c = w['ccf']
    sh = synth.SHAPES[setup['shape']]
    conf = make_ccf.get_ccf_config(np.log(sh['t_lo']), np.log(sh['t_hi']), c['npoints'])
    path = ccf_bank_path(wname)
    if not os.path.exists(path):
        bank = spec_inter.getInterpolator(setup['name'], make_config(w)).bank
        nodes = setup['vec'].T.copy()
        nodes[:, 0] = 10**nodes[:, 0]
        b = make_ccf.build_bank(bank, nodes, conf, every=c['every'], vsinis=c['vsinis'])
        np.savez(path + '.tmp.npz', models=b['models'], params=b['params'],
                 vsinis=np.array(b['vsinis']))
        os.replace(path + '.tmp.npz', path)
    z = np.load(path)
    return conf, z['models'], z['params'], list(z['vsinis'])


def run_gpu_ccf(args):
    """--mode ccf: one step = fitter_ccf.fit_batch over the batch (device preprocessing,
    cuFFT, fused product / window / argmin kernels), host arrays in, first guesses out --
    this path has no resident variant (its inputs are the spectra), so `value` and `e2e`
    are the same measurement."""
    import torch
    from rvspecfit_b200 import _cabi, _dev, fitter_ccf, spec_fit, spec_inter
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    w = WORKLOADS[args.workload]
    cfg = make_config(w)
    B = args.batch
    setups, objects, pars, vel = make_inputs(args.workload, B, 1000)
    for st in setups:
        spec_inter.register_bank(spec_inter.bank_from_setup(st), template_lib='synthetic/')
    conf, models, params, vsinis = ccf_bank(args.workload, setups[0])
    name = setups[0]['name']
    fitter_ccf.register_ccf_bank(name, np.fft.rfft(models, axis=1), np.fft.rfft(models**2, axis=1),
                                 models, params, vsinis, list(setups[0]['parnames']), conf)
    sds = [[spec_fit.SpecData(*a) for a in o] for o in objects]
    L = _cabi.lib()

    def step():
        res = fitter_ccf.fit_batch(sds, cfg, want_proc_spec=False)
        return np.array([[r['best_vel'], r['best_vsini'], r['best_id']] for r in res])
    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    from rvspecfit_b200 import batch_fit
    fitter_ccf.timer = batch_fit.KernelTimer()
    l0 = L.rvs_launch_count()
    io0 = list(_dev.IO_BYTES)
    clk = ClockSampler(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1) / args.steps
    ksum = fitter_ccf.timer.summary()
    fitter_ccf.timer = None
    acc_ms = ksum.get('ccf_accumulate_ms_total', 0.0) / args.steps
    ntempl, npts = len(models), conf['npoints']
    # algorithmic bytes per (object, template): the two template transforms read and the
    # inverse transform's output (SURVEY.md section 8d / DESIGN.md 3.3)
    bytes_step = B * ntempl * 3 * (npts // 2 + 1) * 16
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm = peaks.get('hbm_gbs', 6650.0)
    dv = out[:, 0] - vel
    line = {'metric': 'spectra/sec (CCF first guess)', 'value': B / (ms * 1e-3),
            'unit': 'spectra/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64 / c128', 'data': 'synthetic',
            'config': {'workload': workload_string(args.workload), 'mode': 'ccf',
                       'spectra_per_gpu_per_step': B, 'ccf_templates': ntempl,
                       'ccf_npoints': npts, 'obs_px': len(objects[0][0][1]),
                       'step': 'fitter_ccf.fit_batch: device preprocessing (masks, soft-L1 '
                               'continuum, resampling), rfft of the data, product + inverse '
                               'transform per template, lag window, argmin + parabola'},
            'e2e': {'value': B / (ms * 1e-3), 'unit': 'spectra/s',
                    'h2d_bytes_per_step': int((_dev.IO_BYTES[0] - io0[0]) / args.steps),
                    'd2h_bytes_per_step': int((_dev.IO_BYTES[1] - io0[1]) / args.steps)},
            'gpu_launches': int(L.rvs_launch_count() - l0), 'clocks': clk.stop(t0, t1),
            'roofline': {'bound': 'hbm', 'kernel': 'rvs_ccf_accumulate (rfft of the data, ccf_mult, '
                         'cuFFT Z2D, ccf_gather), CUDA events around every call',
                         'achieved': bytes_step / (acc_ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                         'frac': bytes_step / (acc_ms * 1e-3) / 1e9 / hbm, 'traffic': None,
                         'ms_per_step': acc_ms, 'share_of_step': acc_ms / ms,
                         'whole_step_frac': bytes_step / (ms * 1e-3) / 1e9 / hbm,
                         'algorithmic_bytes_per_object_template': 3 * (npts // 2 + 1) * 16,
                         'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs)' if peaks
                         else 'fallback'},
            'first_guess_rms_kms': float(np.sqrt(np.mean(dv**2))),
            'first_guess_median_abs_kms': float(np.median(np.abs(dv)))}
    if not args.no_cpu:
        line['cpu_baseline'] = cpu_baseline(args)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    saved_stdout = None
    if world > 1:
        # stdout carries exactly one JSON line: anything libraries write to fd 1 while the
        # job runs (NCCL prints its version there when NCCL_DEBUG is set) goes to stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from rvspecfit_b200 import _cabi, _dev, spec_fit, spec_inter, batch_fit, shard
    w = WORKLOADS[args.workload]
    cfg = make_config(w)
    if world > 1 and hasattr(os, 'sched_setaffinity'):
        # every rank (its interpreter and the host threads of its lock-step sets) on its
        # own block of cores, so that the ranks do not migrate over each other
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        mine = cores[local * per:(local + 1) * per] or cores
        try:
            os.sched_setaffinity(0, mine)
        except OSError:
            pass
    if args.total_spectra:
        # strong scaling: a fixed set of spectra, rank r fits its contiguous block
        # (shard.block_range); the set is generated whole so that every N fits the same objects
        lo, hi = shard.block_range(args.total_spectra, rank, world)
        B = hi - lo
        setups, objects, pars, vel = make_inputs(args.workload, args.total_spectra, 1000)
        objects, pars, vel = objects[lo:hi], pars[lo:hi], vel[lo:hi]
    else:
        # weak scaling: every rank owns its own B spectra; the template grid is
        # replicated in each GPU's HBM; no data-path collective
        B = args.batch
        setups, objects, pars, vel = make_inputs(args.workload, B, 1000 + 7919 * rank)
    for st in setups:
        spec_inter.register_bank(spec_inter.bank_from_setup(st), template_lib='synthetic/')
    opts = {'npoly': w['npoly']}
    beval, npt, npo = algorithmic_bytes(setups, objects)
    tp, tv, tvs = trial_points(pars, vel, w['layout'], args.evals, 5 + rank)
    vgrid = np.arange(cfg['min_vel'], cfg['max_vel'], cfg['vel_step0'])
    start = np.tile(np.array([[5500., 3.0, -1.0, 0.2]]), (B, 1))

    resol, nd = [None] * len(setups), 0
    if args.resolution_matrix > 0:
        # diagnostic workload: every spectrum carries a banded resolution matrix (Gaussian
        # line-spread function of that sigma in Angstrom, cut to 11 diagonals like DESI's,
        # desi/desi_fit.py:723-748); the band rows add nd x 8 B per observed pixel
        import scipy.sparse
        for a, arm in enumerate(objects[0]):
            full = scipy.sparse.dia_matrix(
                spec_fit.construct_resol_mat(arm[1], width=args.resolution_matrix).mat)
            keep = np.abs(full.offsets) <= 5
            resol[a] = spec_fit.ResolMatrix(scipy.sparse.dia_matrix(
                (full.data[keep], full.offsets[keep]), shape=full.shape))
        nd = int(keep.sum())
        beval += nd * npo * 8

    def to_specdata():
        return [[spec_fit.SpecData(*a, resolution=r) for a, r in zip(o, resol)] for o in objects]

    timer = batch_fit.KernelTimer()

    starts = [dict(FIT_START) for _ in range(B)]

    def fit_records(res):
        return np.array([[r['vel'], r['vel_err'], r['chisq'], r['vsini']] +
                         [r['param'][k] for k in synth.PARNAMES] for r in res])

    def hot_path(eng):
        """The step body on a ready engine; returns a small result array."""
        if args.mode == 'fit':
            return fit_records(batch_fit.process_batch(None, starts, config=cfg, options=opts,
                                                       engine=eng, timer=timer,
                                                       groups=args.groups or None))
        return batch_fit.scan_and_evaluate(eng, start, vgrid, tp, tv, tvs, timer=timer,
                                           groups=args.groups or 2)

    def step_resident(eng):
        return hot_path(eng)

    def glaunch():
        # kernels launched through CUDA-graph replays (the library counter only sees
        # direct launches)
        return spec_fit.GRAPH_LAUNCHES[0]

    host_objects = to_specdata()    # host containers of the input arrays (reference SpecData)

    e2e_parts = {'engine_load_s': [], 'hot_path_s': []}

    e2e_eng = []

    def step_e2e():
        # host arrays -> pinned staging -> HBM (spectra, errors), derived products on the
        # device, then the hot path and the D2H of its results.  The engine is the
        # persistent one of a survey driver: built by the first (warm-up) step, it receives
        # every later step's spectra through LikelihoodEngine.reload (same instrument, new
        # exposure), which keeps device addresses and captured graphs
        t0 = time.time()
        if not e2e_eng:
            e2e_eng.append(spec_fit.LikelihoodEngine(host_objects, cfg, opts))
        else:
            e2e_eng[0].reload(host_objects)
        eng = e2e_eng[0]
        t1 = time.time()
        rec = hot_path(eng)                                          # D2H of the results
        e2e_parts['engine_load_s'].append(t1 - t0)
        e2e_parts['hot_path_s'].append(time.time() - t1)
        e2e_recs.append(rec)
        return rec

    e2e_recs = []

    def e2e_finish():
        # the one collective of the path: the fixed-size result records of all ranks,
        # gathered ONCE after the last step (inside the timed region) -- the ranks never
        # wait for each other while they fit
        if world > 1 and e2e_recs:
            allrec = np.concatenate(e2e_recs)
            return shard.gather_records(allrec, len(allrec) * world)
        return None

    eng = spec_fit.LikelihoodEngine(host_objects, cfg, opts)
    L = _cabi.lib()
    if args.timeline:
        # diagnostic: start/end of every kernel of a few evaluation calls, concurrent
        # streams as in the bench
        import ctypes
        tp2, tv2, tvs2 = tp[:args.timeline], tv[:args.timeline], tvs[:args.timeline]
        batch_fit.scan_and_evaluate(eng, start, vgrid[:8], tp2, tv2, tvs2, groups=args.groups or 2)
        L.rvs_profile_enable(1)
        batch_fit.scan_and_evaluate(eng, start, vgrid[:8], tp2, tv2, tvs2, groups=args.groups or 2)
        buf = (ctypes.c_double * (3 * 4096))()
        n = L.rvs_profile_timeline(ctypes.cast(buf, ctypes.c_void_p), 4096)
        L.rvs_profile_enable(0)
        names = ['locate', 'prep', 'chunk', 'gram', 'solve', 'resid']
        for i in range(n):
            print(f'{names[int(buf[3 * i])]:8s} {1e3 * buf[3 * i + 1]:10.1f} {1e3 * buf[3 * i + 2]:10.1f}'
                  f'  dur {1e3 * (buf[3 * i + 2] - buf[3 * i + 1]):8.1f} us')
        return
    if args.stage_profile:
        # diagnostic run (not a bench value): arms serialised on one stream, CUDA
        # events around every kernel of the evaluation call
        import ctypes
        eng.serial_arms = True
        for _ in range(2):
            step_resident(eng)
        L.rvs_profile_enable(1)
        step_resident(eng)
        ms = (ctypes.c_double * 6)()
        n = (ctypes.c_int64 * 6)()
        L.rvs_profile_read(ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(n, ctypes.c_void_p), 6)
        L.rvs_profile_enable(0)
        names = ['locate', 'prep', 'chunk', 'gram_mma', 'gram_solve', 'resid_mma']
        out = {k: dict(launches=int(c), us_per_launch=1e3 * t / max(1, c), ms_total=t)
               for k, t, c in zip(names, ms, n)}
        out['note'] = ('arms serialised on one stream; warm caches; per-launch = one arm, '
                       f'{B // max(1, args.groups or 2)} items')
        if rank == 0:
            print(json.dumps({'stage_profile': out}))
        return

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps, nwarm, sample_clocks=False, finish=None):
        for _ in range(nwarm):
            fn()
        timer.reset()
        e2e_recs.clear()
        barrier()
        io0 = list(_dev.IO_BYTES)
        l0 = L.rvs_launch_count() + glaunch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk = ClockSampler(local) if sample_clocks else None
        t0 = time.time()
        e0.record()
        for _ in range(nsteps):
            out = fn()
        if finish is not None:
            finish()
        e1.record()
        barrier()
        t1 = time.time()
        timed.io = [(b - a) / nsteps for a, b in zip(io0, _dev.IO_BYTES)]
        ms = e0.elapsed_time(e1)
        clocks = clk.stop(t0, t1) if clk else None
        if world > 1:
            t = torch.tensor([ms], device='cuda', dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, L.rvs_launch_count() + glaunch() - l0, clocks, out, timer.summary()

    ms, launches, clocks, out, ksum = timed(lambda: step_resident(eng), args.steps, args.warmup,
                                            sample_clocks=True)
    # end to end: every requested step again, from host arrays (the first warm-up step
    # builds the persistent engine)
    n_e2e = args.steps
    if args.no_e2e:         # profiler runs only: the line then carries no end-to-end figure
        ms_e2e, timed.io = float('nan'), [0, 0]
    else:
        ms_e2e, _, _, _, _ = timed(step_e2e, n_e2e, max(2, min(args.warmup, 3)),
                                   finish=e2e_finish)
    e2e_eng.clear()
    # bytes per timed step, counted where the copies are made (_dev.IO_BYTES): flux and
    # error of every spectrum through LikelihoodEngine.reload (+ band rows of resolution
    # matrices), the velocity grids / parameters / arm indices of every scan and evaluation
    # call, and back the chi-squares, flags, scan statistics and best-fit models; whole job
    h2d = int(timed.io[0]) * world
    d2h = int(timed.io[1]) * world

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else \
        'fallback 6650 GB/s (B200_PROFILING.md)'
    traffic = None
    try:    # DRAM bytes of one evaluation call from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r2m_traffic.json')))
        if args.workload == 'desi':
            traffic = tj['bytes_per_call'] / tj['items_per_call']    # per evaluation
    except Exception:
        pass
    roof = None
    kern_txt = ('rvs_locate_grid + rvs_chisq_fused (prep_kernel, chunk_kernel [TMA gather -> exp -> '
                'broadening -> spline -> resampling], gram_mma/gram_solve/resid_mma kernels) of the '
                '3 arms')
    if ksum.get('eval_phase_ms_total'):
        # the optimiser-phase evaluations of the timed region: CUDA events bracket the
        # whole phase (in-flight evaluations of different object groups overlap on the
        # GPU, so per-call durations do not add up); bytes = SURVEY.md 8d's
        # per-evaluation figure x evaluations
        evals = ksum['eval_phase_items_per_launch'] * ksum['eval_phase_launches']
        ach = beval * evals / (ksum['eval_phase_ms_total'] * 1e-3) / 1e9
        ncall = max(1, ksum.get('fused_eval_launches', 1))
        roof = dict(bound='hbm', kernel=kern_txt, achieved=ach, peak=hbm_peak, unit='GB/s',
                    frac=ach / hbm_peak,
                    traffic=None if traffic is None else traffic * (B // max(1, args.groups or 2)),
                    traffic_bytes_per_eval=traffic,
                    traffic_source='static: ncu --set full capture of a 2048-item call, '
                                   'profiles/r2m_traffic.json (not re-measured in this run)',
                    peak_source=peak_src,
                    algorithmic_bytes_per_eval=beval, evals_timed=evals,
                    ms_total=ksum['eval_phase_ms_total'],
                    ms_per_call=ksum['eval_phase_ms_total'] / ncall,
                    items_per_call=ksum.get('fused_eval_items_per_launch'),
                    timing='CUDA events around the evaluation phase of every step')
    elif ksum.get('fused_eval_ms_total'):
        # complete-fit mode: CUDA events around every evaluation call, on the stream its
        # arms fork from and join to.  The lock-step sets of a step keep several calls in
        # flight at once, so the time the evaluation kernels had the device is the UNION of
        # the calls' [start, end] intervals, not their sum; RV scans of other sets that run
        # inside those intervals are not subtracted (their time counts against the figure)
        evals = ksum['fused_eval_items_per_launch'] * ksum['fused_eval_launches']
        ach = beval * evals / (ksum['fused_eval_ms_busy'] * 1e-3) / 1e9
        roof = dict(bound='hbm', kernel=kern_txt, achieved=ach, peak=hbm_peak, unit='GB/s',
                    frac=ach / hbm_peak,
                    traffic=None if traffic is None else traffic * ksum['fused_eval_items_per_launch'],
                    traffic_bytes_per_eval=traffic,
                    traffic_source='static: ncu --set full capture of a 2048-item call, '
                                   'profiles/r2m_traffic.json (not re-measured in this run)',
                    peak_source=peak_src, algorithmic_bytes_per_eval=beval,
                    evals_timed=evals, ms_busy=ksum['fused_eval_ms_busy'],
                    ms_sum_of_calls=ksum['fused_eval_ms_total'],
                    ms_per_call=ksum['fused_eval_ms_per_launch'],
                    items_per_call=ksum['fused_eval_items_per_launch'],
                    share_of_step=ksum['fused_eval_ms_busy'] / ms,
                    timing='CUDA events around every evaluation call; busy time = union '
                           'of the call intervals over the timed steps')
    nspec_total = args.total_spectra or B * world
    per_step = ms / args.steps
    fit = args.mode == 'fit'
    step_txt = ('per spectrum: the complete vel_fit.process fit (RV-grid scan, Nelder-Mead, '
                'BFGS, RV refinement scans, model, Hessian) through batch_fit.process_batch'
                if fit else
                'per spectrum: 1 RV-grid scan + fit_evals_per_spectrum template-changing chi2 '
                'evaluations (count of one reference process() call, SURVEY.md 8d)')
    line = {
        'metric': 'spectra/sec (RV-grid chi2 + fit)' if fit else
                  'spectra/sec (RV-grid chi2 scan + fit evaluations)',
        'value': nspec_total / (per_step * 1e-3), 'unit': 'spectra/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': per_step,
        'higher_is_better': True, 'scaling': 'strong' if args.total_spectra else 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_string(args.workload),
                   'obs_px': npo, 'template_px': npt,
                   'spectra_per_gpu_per_step': B, 'rv_trials': len(vgrid),
                   'fit_evals_per_spectrum': (eng.n_eval / (args.steps + args.warmup) / B
                                              if fit else args.evals),
                   'lockstep_groups': (args.groups if args.groups else 'auto'),
                   'fit_driver': {'sets': args.groups or batch_fit.FIT_MAX_GROUPS,
                                  'set_sizes': batch_fit.FIT_SPLIT.get(
                                      args.groups or batch_fit.FIT_MAX_GROUPS),
                                  'handover_min_objects': batch_fit.PEEL_MIN if batch_fit.PEEL else None,
                                  'speculate_below': batch_fit.SPECULATE_BELOW,
                                  'native_round_loop': bool(batch_fit.NATIVE_DRIVE),
                                  'native_bfgs': bool(batch_fit.NATIVE_BFGS)},
                   'mode': args.mode, 'step': step_txt, 'resolution_matrix_diagonals': nd,
                   'l2': 'template grid (>=0.7 GB per arm) is larger than L2; rows gathered '
                         'at random per evaluation',
                   'parallelism': f'spectra sharded over {world} GPU(s), grid replicated'},
        'chisq_evals_per_s': (eng.n_eval / (args.steps + args.warmup) * world if fit else
                              nspec_total * (args.evals + len(vgrid))) / (per_step * 1e-3),
        'e2e': {'value': nspec_total / (ms_e2e / n_e2e * 1e-3), 'unit': 'spectra/s',
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
        'e2e_host_seconds_per_step': {k: float(np.mean(v[-n_e2e:])) for k, v in e2e_parts.items()},
        'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof,
        'fit_phase_seconds': getattr(batch_fit.process_batch, 'last_phase_seconds', None),
        'kernels': ksum,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        line['cpu_baseline'] = cpu_baseline(args, bounded=True)
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


# ------------------------------------------------------------------ CPU arm
def _reference_available():
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_loader
    return ref_loader.available()


def _cpu_worker(job):
    """One object through the CPU implementation of the path -- the reference package
    itself when baseline/_ref holds it (baseline/install_ref.py), else the oracle port:
    the same step body as the GPU arm."""
    wname, seed, idx, nevals, nscan = job
    os.environ['OMP_NUM_THREADS'] = '1'
    w = WORKLOADS[wname]
    cfg = make_config(w)
    key = (wname, seed)
    if _cpu_worker.cache.get('key') != key:
        setups, objects, pars, vel = make_inputs(wname, _cpu_worker.nspec, seed, mmap_grids=True)
        if _reference_available():
            import ref_loader
            R = ref_loader.load()
            for st in setups:
                ref_loader.inject_grid(R, st)
            api = dict(SpecData=R.spec_fit.SpecData, process=R.vel_fit.process,
                       find_best=R.spec_fit.find_best, get_chisq=R.spec_fit.get_chisq,
                       config=R.utils.freezeDict(cfg), kind='reference')
        else:
            sys.path.insert(0, os.path.join(ROOT, 'oracle'))
            import oracle
            for st in setups:
                oracle.register_setup(st)
            api = dict(SpecData=oracle.SpecData, process=oracle.process,
                       find_best=lambda sd, vg, pl, rot_params=None, **kw:
                       oracle.find_best(sd, vg, pl, rot=rot_params, **kw),
                       get_chisq=oracle.get_chisq, config=cfg, kind='port')
        _cpu_worker.cache = dict(key=key, objects=objects, pars=pars, vel=vel, api=api)
    c = _cpu_worker.cache
    api = c['api']
    cfg = api['config']
    sd = [api['SpecData'](a[0], a[1], a[2], a[3], badmask=a[4]) for a in c['objects'][idx]]
    if nevals == -2:     # --mode ccf: the cross-correlation first guess
        if 'ccf' not in c:
            z = np.load(ccf_bank_path(wname))
            sh = synth.SHAPES[WORKLOADS[wname]['arms'][0]]
            if api['kind'] == 'reference':
                import ref_loader
                R = ref_loader.load()
                conf = R.make_ccf.get_ccf_config(logl0=np.log(sh['t_lo']), logl1=np.log(sh['t_hi']),
                                                 npoints=w['ccf']['npoints'])
                CC = R.fitter_ccf.CCFCache
                nm = w['arms'][0]
                CC.ccfs[nm] = np.fft.rfft(z['models'], axis=1)
                CC.ccf2s[nm] = np.fft.rfft(z['models']**2, axis=1)
                CC.ccf_models[nm] = z['models']
                CC.ccf_info[nm] = dict(params=z['params'], ccfconf=conf, vsinis=list(z['vsinis']),
                                       parnames=list(synth.PARNAMES))
                c['ccf'] = lambda s: R.fitter_ccf.fit(s, cfg)
            else:
                import oracle
                conf = oracle.ccf_config(np.log(sh['t_lo']), np.log(sh['t_hi']), w['ccf']['npoints'])
                bank = {w['arms'][0]: dict(
                    fft=np.fft.rfft(z['models'], axis=1), fft2=np.fft.rfft(z['models']**2, axis=1),
                    models=z['models'], params=z['params'], vsinis=list(z['vsinis']),
                    parnames=list(synth.PARNAMES), ccfconf=conf)}
                c['ccf'] = lambda s: oracle.ccf_fit(s, cfg, bank)
        t0 = time.time()
        r = c['ccf'](sd)
        return time.time() - t0, r['best_vel'], api['kind']
    opts = {'npoly': w['npoly']}
    tp, tv, tvs = trial_points(c['pars'], c['vel'], w['layout'], max(nevals, 0), 5)
    vgrid = np.arange(w['min_vel'], w['max_vel'], 5)[:nscan]
    t0 = time.time()
    if nevals < 0:       # --mode fit: the complete fit
        r = api['process'](sd, dict(FIT_START), fixParam=[], options=opts, config=cfg)
        return time.time() - t0, r['chisq'], api['kind']
    fb = api['find_best'](sd, vgrid, [(5500., 3.0, -1.0, 0.2)], rot_params=None, options=opts,
                          config=cfg)
    acc = fb['best_chi']
    for e in range(nevals):
        acc += api['get_chisq'](sd, tv[e, idx], tuple(tp[e, idx]), (tvs[e, idx],),
                                options=opts, config=cfg)
    return time.time() - t0, acc, api['kind']


_cpu_worker.cache = {}
_cpu_worker.nspec = 0


def _cpu_init(nspec):
    _cpu_worker.nspec = nspec
    import logging
    import warnings
    warnings.simplefilter('ignore')
    logging.disable(logging.WARNING)     # the reference logs a line per ill-conditioned Hessian


def cpu_baseline(args, bounded=True):
    """The reference's own CPU implementation (baseline/_ref; the oracle port if that
    is absent) on all host cores, one object per task in a spawn pool (the reference's
    own pattern, desi_fit.py:1475-1479), on a bounded sample of the same workload."""
    import concurrent.futures as cf
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    w = WORKLOADS[args.workload]
    vg = np.arange(w['min_vel'], w['max_vel'], 5)
    frac = args.cpu_fraction
    nevals = max(1, int(args.evals * frac))
    nscan = max(3, int(len(vg) * frac))
    nobj = 2 * cores
    if args.mode == 'fit':
        nevals, nscan, nobj = -1, len(vg), cores
    if args.mode == 'ccf':
        nevals, nscan, nobj = -2, 0, 4 * cores
        if not os.path.exists(ccf_bank_path(args.workload)):
            raise RuntimeError('the CCF bank of the workload has not been built: run the GPU '
                               'arm of --mode ccf once on this box first')
    for k, a in enumerate(w['arms']):       # template rows shared through the page cache
        path = grid_cache_path(args.workload, a)
        if not os.path.exists(path):
            st = synth.make_setup(a, w['layout'], seed=21 + k)
            np.save(path + '.tmp.npy', st['dats'])
            os.replace(path + '.tmp.npy', path)
    jobs = [(args.workload, 1000, i, nevals, nscan) for i in range(nobj)]
    t0 = time.time()
    with cf.ProcessPoolExecutor(cores, mp_context=mp.get_context('spawn'),
                                initializer=_cpu_init, initargs=(nobj,)) as ex:
        list(ex.map(_cpu_worker, [(j[0], j[1], j[2], -2 if args.mode == 'ccf' else 1, 3)
                                  for j in jobs[:cores]]))       # warm-up: per-worker banks
        t1 = time.time()
        res = list(ex.map(_cpu_worker, jobs))
        t2 = time.time()
    wall = t2 - t1
    kind = res[0][2]
    impl = ('the reference package (baseline/_ref, vel_fit.process)' if kind == 'reference'
            else 'the oracle port (oracle.process)')
    if args.mode == 'ccf':
        return dict(value=nobj / wall, unit='spectra/s', cores=cores, kind=kind,
                    sample=f'{nobj} first guesses by '
                           f'{"the reference package (fitter_ccf.fit)" if kind == "reference" else "the oracle port (oracle.ccf_fit)"}'
                           f' on {cores} processes, one object per task; setup {t1 - t0:.1f}s '
                           'excluded', wall_s=wall,
                    per_object_s=float(np.mean([r[0] for r in res])))
    if args.mode == 'fit':
        return dict(value=nobj / wall, unit='spectra/s', cores=cores, kind=kind,
                    sample=f'{nobj} complete fits by {impl}: scan, Nelder-Mead, BFGS, '
                           f'refinement, Hessian; {cores} processes, one object per task; '
                           f'setup {t1 - t0:.1f}s excluded', wall_s=wall,
                    per_object_s=float(np.mean([r[0] for r in res])))
    scale = (args.evals + len(vg)) / (nevals + nscan)
    return dict(value=nobj / (wall * scale), unit='spectra/s', cores=cores, kind=kind,
                sample=f'{nobj} spectra x ({nscan} RV trials + {nevals} evaluations) on {cores} '
                       f'processes, scaled x{scale:.1f} to the full per-spectrum count; '
                       f'setup {t1 - t0:.1f}s excluded', wall_s=wall,
                per_object_s=float(np.mean([r[0] for r in res])) * scale)


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    vals = []
    for _ in range(args.warmup + args.steps):
        vals.append(cpu_baseline(args))
    cb = vals[-1]
    v = float(np.mean([x['value'] for x in vals[args.warmup:]])) if args.steps else cb['value']
    cb['value'] = v
    setups = [a for a in w['arms']]
    line = {'impl': 'reference',
            'metric': {'fit': 'spectra/sec (RV-grid chi2 + fit)',
                       'ccf': 'spectra/sec (CCF first guess)'}.get(
                           args.mode, 'spectra/sec (RV-grid chi2 scan + fit evaluations)'),
            'value': v, 'unit': 'spectra/s', 'n_gpus': int(os.environ.get('WORLD_SIZE', 1)),
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': cb['wall_s'] * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': workload_string(args.workload), 'mode': args.mode},
            'cpu_baseline': cb,
            'e2e': {'value': v, 'unit': 'spectra/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='desi', choices=list(WORKLOADS))
    ap.add_argument('--batch', type=int, default=4096, help='spectra per GPU per step')
    ap.add_argument('--total-spectra', type=int, default=0,
                    help='strong scaling: this many spectra in total, split over the ranks in '
                         'contiguous blocks (default: weak scaling, --batch per GPU)')
    ap.add_argument('--evals', type=int, default=EVALS_PER_FIT)
    ap.add_argument('--cpu-fraction', type=float, default=1.0,
                    help='fraction of the per-spectrum evaluations the CPU sample runs')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true',
                    help='diagnostic (profiler runs): skip the end-to-end leg')
    ap.add_argument('--resolution-matrix', type=float, default=0.0, metavar='SIGMA_A',
                    help='diagnostic: attach an 11-diagonal resolution matrix (Gaussian of this '
                         'sigma in Angstrom) to every spectrum (SURVEY.md 8 row f4)')
    ap.add_argument('--timeline', type=int, default=0,
                    help='diagnostic: kernel start/end times of this many evaluation rounds')
    ap.add_argument('--stage-profile', action='store_true',
                    help='diagnostic: CUDA-event time of every kernel of the evaluation call')
    ap.add_argument('--mode', default='fit', choices=['fit', 'proxy', 'ccf'],
                    help='fit: complete vel_fit.process fits (the headline); proxy: one RV scan '
                         '+ a fixed count of evaluations at pre-generated points (kernel study)')
    ap.add_argument('--groups', type=int, default=0,
                    help='independent lock-step sets of objects in flight (0: automatic)')
    args = ap.parse_args()
    if args.stage_profile:
        # per-kernel times are only meaningful when nothing else runs beside the kernel:
        # one group, arms serialised, no helper stream
        os.environ['RVS_NO_AUX'] = '1'
        args.groups = 1
    if args.impl == 'reference':
        run_reference(args)
    elif args.mode == 'ccf':
        run_gpu_ccf(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
