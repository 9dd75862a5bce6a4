"""Cross-correlation first guess on the GPU.

Mirror of the reference's fitter_ccf.py (paths under
/root/reference/py/rvspecfit/): `fit(specdata, config)` keeps its signature,
returned keys and error behaviour (fitter_ccf.py:62-253); `fit_batch` is the
batched sibling that runs many objects through the same launches.  The host
preprocesses each arm (make_ccf.preprocess_data, row f3 of SURVEY.md section 8)
and derives the lag -> velocity-grid table; the transforms (cuFFT), the spectral
products, the window interpolation, the sum over arms and the argmin / parabola
run in librvs_b200.so (rvs_ccf_accumulate, rvs_ccf_best).
"""
import ctypes
import logging

import numpy as np

from . import _cabi, _dev, make_ccf
from .spec_fit import SpecData

# device workspace the batched path may use per call (bytes)
WORKSPACE_BYTES = 4 << 30
# cap on the [objects, templates, velocities] chi-square block held at once
CHISQ_BLOCK_BYTES = 2 << 30


class CcfBank:
    """CCF templates of one spectral setup, resident in HBM (the reference's
    ccfdat_<setup>.npz / ccfmod_<setup>.npy / ccf_<setup>.h5 products,
    make_ccf.py:483-493, fitter_ccf.py:39-56)."""

    def __init__(self, name, fft, fft2, models, params, vsinis, parnames, ccfconf):
        self.name = name
        self.models = np.asarray(models)
        self.params = np.asarray(params)
        self.vsinis = np.asarray(vsinis)
        self.parnames = list(parnames)
        self.ccfconf = dict(ccfconf)
        fft = np.ascontiguousarray(fft, dtype=np.complex128)
        fft2 = np.ascontiguousarray(fft2, dtype=np.complex128)
        self.ntempl, nfreq = fft.shape
        self.npoints = int(ccfconf['npoints'])
        if nfreq != self.npoints // 2 + 1 or fft2.shape != fft.shape:
            raise ValueError(f'CCF bank {name}: transform shape {fft.shape} does not match '
                             f'npoints={self.npoints}')
        self.d_fft = _dev.upload(fft.view(np.float64), np.float64)
        self.d_fft2 = _dev.upload(fft2.view(np.float64), np.float64)
        self._tables = {}

    @property
    def velstep(self):
        c = self.ccfconf
        return (np.exp((c['logl1'] - c['logl0']) / c['npoints']) - 1) * make_ccf.C_CCF

    def lag_table(self, maxvel, vel_grid):
        """Which CCF pixels bracket every point of the common velocity grid
        (fitter_ccf.py:132-158 followed by scipy interp1d's index logic)."""
        key = (float(maxvel), len(vel_grid))
        if key not in self._tables:
            step, n = self.velstep, self.npoints
            off = n // 2
            lag_vel = -((np.arange(n) + off) % n - off) * step
            keep = np.abs(lag_vel) < (maxvel + step)
            assert keep.sum() % 2 == 1
            subind = np.roll(np.nonzero(keep)[0], keep.sum() // 2)[::-1]
            x = lag_vel[subind]
            if not np.all(np.diff(x) > 0):
                raise RuntimeError('Velocity grid for CCF interpolation is invalid')
            if vel_grid[0] < x[0] or vel_grid[-1] > x[-1]:
                raise ValueError('A value in x_new is outside the interpolation range.')
            hi = np.clip(np.searchsorted(x, vel_grid), 1, len(x) - 1)
            lo = hi - 1
            tab = dict(lo=_dev.upload(subind[lo], np.int32), hi=_dev.upload(subind[hi], np.int32),
                       dxn=_dev.upload(vel_grid - x[lo], np.float64),
                       dx=_dev.upload(x[hi] - x[lo], np.float64))
            arm = _cabi.CcfArm()
            arm.d_fft, arm.d_fft2 = self.d_fft.data_ptr(), self.d_fft2.data_ptr()
            arm.d_lo, arm.d_hi = tab['lo'].data_ptr(), tab['hi'].data_ptr()
            arm.d_dxn, arm.d_dx = tab['dxn'].data_ptr(), tab['dx'].data_ptr()
            arm.npoints, arm.ntempl = self.npoints, self.ntempl
            arm.continuum, arm.nvel = int(bool(self.ccfconf['continuum'])), len(vel_grid)
            tab['arm'] = arm
            self._tables[key] = tab
        return self._tables[key]


class CCFCache:
    """Registry of CCF banks (reference fitter_ccf.py:13-18)."""
    banks = {}


def register_ccf_bank(name, fft, fft2, models, params, vsinis, parnames, ccfconf):
    """Put an in-memory CCF bank into the registry `get_ccf_info` serves."""
    CCFCache.banks[name] = CcfBank(name, fft, fft2, models, params, vsinis, parnames, ccfconf)
    return CCFCache.banks[name]


def get_ccf_info(spec_setup, config):
    """reference fitter_ccf.py:21-59: registered banks are served as is;
    otherwise the reference's on-disk products are loaded (needs h5py)."""
    if spec_setup not in CCFCache.banks:
        from . import bank_io
        CCFCache.banks[spec_setup] = bank_io.load_ccf_bank(spec_setup, config)
    return CCFCache.banks[spec_setup]


def _velocity_grid(config):
    maxvel = config.get('max_vel') or 1000
    nvel = 2 * int(maxvel * 1. / (config.get('vel_step0') or 2)) + 1
    return maxvel, np.linspace(-maxvel, maxvel, nvel)


_ws = {}


def _workspace(nbytes):
    torch = _dev.torch_mod()
    key = torch.cuda.current_device()
    if key not in _ws or _ws[key].numel() * 8 < nbytes:
        _ws[key] = None
        _ws[key] = _dev.empty((int(nbytes + 7) // 8,), np.float64)
    return _ws[key]


def _stack_rows(pairs):
    """[(proc_spec, proc_ivar)] rows, host arrays or device rows -> two (n, npoints)
    device tensors."""
    torch = _dev.torch_mod()
    if all(isinstance(p[0], np.ndarray) for p in pairs):
        return (_dev.upload(np.stack([p[0] for p in pairs]), np.float64),
                _dev.upload(np.stack([p[1] for p in pairs]), np.float64))
    cols = []
    for k in (0, 1):
        cols.append(torch.stack([p[k] if not isinstance(p[k], np.ndarray)
                                 else _dev.upload(p[k], np.float64) for p in pairs]))
    return cols[0].contiguous(), cols[1].contiguous()


# optional CUDA-event timer of the accumulate stage (batch_fit.KernelTimer; bench.py)
timer = None


def fit_batch(objects, config, preprocessed=None, raise_errors=True, workers=None,
              preprocess='device', want_proc_spec=True):
    """fitter_ccf.fit for many objects.  objects: list of lists of SpecData.
    preprocessed: optional list (per object) of dicts setup -> (proc_spec,
    proc_ivar) that bypasses the preprocessing.  preprocess: 'device' -- masks, gap
    bridging, continuum fit and resampling on the GPU for the spectra of every pixel grid
    in one launch (make_ccf.DevicePrep; spectra it does not take go to the host route);
    'host' -- the reference's own steps (make_ccf.preprocess_data) in a pool of host
    processes.  want_proc_spec=False leaves 'proc_spec' out of the results (saves the
    download in throughput runs).  Returns a list of the reference's result dictionaries
    (plus 'best_id')."""
    L = _cabi.lib()
    objects = [[o] if isinstance(o, SpecData) else list(o) for o in objects]
    nobj = len(objects)
    if nobj == 0:
        return []
    maxvel, vel_grid = _velocity_grid(config)
    nvel = len(vel_grid)
    banks = {}
    checked = set()     # pairs of setups whose banks were compared
    for o in objects:
        ref = banks.get(o[0].name)
        if ref is None:
            ref = banks[o[0].name] = get_ccf_info(o[0].name, config)
        for sd in o:
            b = banks.get(sd.name)
            if b is None:
                b = banks[sd.name] = get_ccf_info(sd.name, config)
            if b is ref or (ref.name, b.name) in checked:
                continue
            if (ref.parnames != b.parnames or not np.array_equal(ref.params, b.params)
                    or not np.array_equal(ref.vsinis, b.vsinis)):
                raise RuntimeError('The parameters of the CCF templates do not match')
            if b.ntempl != ref.ntempl:
                raise RuntimeError('CCF template counts are inconsistent across setups')
            checked.add((ref.name, b.name))
    ntempl = next(iter(banks.values())).ntempl
    # host preprocessing (row f3): proc[i][a] = (proc_spec, proc_ivar); the continuum fits
    # of a batch run in a pool of host processes (make_ccf.preprocess_many)
    jobs, where = [], []
    proc = [[None] * len(o) for o in objects]
    dev_groups = {}     # (setup, pixel grid) -> [(i, a)] for the device route
    for i, o in enumerate(objects):
        for a, sd in enumerate(o):
            if preprocessed is not None:
                proc[i][a] = tuple(np.asarray(_, dtype=np.float64)
                                   for _ in preprocessed[i][sd.name])
            elif preprocess == 'device' and np.isfinite(sd.spec).all() and \
                    make_ccf.device_prep(sd.lam, sd.gridkey, banks[sd.name].ccfconf).ok:
                dev_groups.setdefault((sd.name, sd.gridkey), []).append((i, a))
            else:
                jobs.append((sd.lam, sd.spec, sd.espec, sd.badmask, banks[sd.name].ccfconf))
                where.append((i, a))
    for (i, a), res in zip(where, make_ccf.preprocess_many(jobs, workers)):
        proc[i][a] = res
    for (name, gkey), members in dev_groups.items():
        first = objects[members[0][0]][members[0][1]]
        prep = make_ccf.device_prep(first.lam, gkey, banks[name].ccfconf)
        sds = [objects[i][a] for i, a in members]
        d_ps, d_pi = prep([s.spec for s in sds], [s.espec for s in sds],
                          [s.badmask for s in sds])
        for r, (i, a) in enumerate(members):
            proc[i][a] = (d_ps[r], d_pi[r])          # device rows
    d_vg = _dev.upload(vel_grid, np.float64)
    block = max(1, int(CHISQ_BLOCK_BYTES // (ntempl * nvel * 8)))
    results = [None] * nobj
    torch = _dev.torch_mod()

    def enqueue(i0):
        """Device work of the block of objects from i0: correlation of every arm, best
        template and lag, results on their way to pinned host memory behind an event."""
        ids = range(i0, min(nobj, i0 + block))
        nrow = len(ids)
        d_chisq = _dev.zeros((nrow, ntempl, nvel), np.float64)
        d_sse = _dev.zeros((nrow,), np.float64)
        # arms in each object's own order, so that the sum over arms is taken in
        # the order the reference takes it (fitter_ccf.py:184-206)
        for a in range(max(len(objects[i]) for i in ids)):
            groups = {}
            for r, i in enumerate(ids):
                if a < len(objects[i]):
                    groups.setdefault(objects[i][a].name, []).append(r)
            for name, rows in groups.items():
                bank = banks[name]
                tab = bank.lag_table(maxvel, vel_grid)
                d_ps, d_pi = _stack_rows([proc[i0 + r][a] for r in rows])
                d_row = _dev.upload(np.asarray(rows), np.int32)
                arm = tab['arm']
                need = L.rvs_ccf_workspace(ctypes.byref(arm), len(rows))
                nbytes = min(need, max(WORKSPACE_BYTES, L.rvs_ccf_workspace(ctypes.byref(arm), 1)))
                ws = _workspace(nbytes)
                t0 = timer.start() if timer is not None else None
                rc = L.rvs_ccf_accumulate(ctypes.byref(arm), _dev.ptr(d_ps), _dev.ptr(d_pi),
                                          len(rows), _dev.ptr(d_row), _dev.ptr(d_chisq),
                                          _dev.ptr(d_sse), _dev.ptr(ws), ws.numel() * 8,
                                          _dev.stream())
                _cabi.check(rc, 'rvs_ccf_accumulate')
                if t0 is not None:
                    timer.stop('ccf_accumulate', t0, len(rows) * ntempl)
        d_out = _dev.empty((nrow, 8), np.float64)
        d_best = _dev.empty((nrow, nvel), np.float64)
        rc = L.rvs_ccf_best(_dev.ptr(d_chisq), _dev.ptr(d_sse), _dev.ptr(d_vg), nrow, ntempl, nvel,
                            _dev.ptr(d_out), _dev.ptr(d_best), _dev.stream())
        _cabi.check(rc, 'rvs_ccf_best')
        h_out = torch.empty((nrow, 8), dtype=torch.float64, pin_memory=True)
        h_best = torch.empty((nrow, nvel), dtype=torch.float64, pin_memory=True)
        h_out.copy_(d_out, non_blocking=True)
        h_best.copy_(d_best, non_blocking=True)
        _dev.IO_BYTES[1] += (h_out.numel() + h_best.numel()) * 8
        done = torch.cuda.Event()
        done.record()
        return ids, h_out, h_best, done, (d_out, d_best, d_chisq)

    def rolled(model, shift):
        """np.roll(model, shift) (fitter_ccf.py:240-244) without its bookkeeping."""
        n = len(model)
        s_ = shift % n
        out = np.empty_like(model)
        out[s_:] = model[:n - s_]
        out[:s_] = model[n - s_:]
        return out

    def finish(ids, h_out, h_best, done, _keep):
        """Result dictionaries of a block (host work: runs under the next block's kernels)."""
        done.synchronize()
        out = h_out.numpy()
        best = h_best.numpy().copy()        # rows outlive the pinned buffer
        for r, i in enumerate(ids):
            if not out[r, 4]:
                logging.error('Cross-correlation failed')
                if raise_errors:
                    raise RuntimeError('Cross-correlation step failed')
                continue
            bid, bvel = int(out[r, 0]), float(out[r, 2])
            ref = banks[objects[i][0].name]
            results[i] = dict(
                best_par=dict(zip(ref.parnames, ref.params[bid])), best_vel=bvel,
                best_ccf=best[r], best_vsini=ref.vsinis[bid],
                best_model={sd.name: rolled(banks[sd.name].models[bid],
                                            int(bvel / banks[sd.name].velstep))
                            for sd in objects[i]},
                vel_grid=vel_grid, best_id=bid)
            if want_proc_spec:
                results[i]['proc_spec'] = {
                    sd.name: (proc[i][a][0] if isinstance(proc[i][a][0], np.ndarray)
                              else _dev.download(proc[i][a][0]))
                    for a, sd in enumerate(objects[i])}

    # blocks in a two-deep pipeline: the host assembles the results of one block while the
    # device correlates the next
    if nobj >= 256:
        block = min(block, -(-nobj // 4))
    pending = None
    for i0 in range(0, nobj, block):
        ctx = enqueue(i0)
        if pending is not None:
            finish(*pending)
        pending = ctx
    if pending is not None:
        finish(*pending)
    return results


def fit(specdata, config):
    """Cross-correlate the data with the template bank: reference
    fitter_ccf.py:62-253, same arguments and returned keys."""
    if isinstance(specdata, SpecData):
        specdata = [specdata]
    return fit_batch([list(specdata)], config, preprocess='host')[0]
