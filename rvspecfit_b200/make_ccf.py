"""Host-side preparation around the CCF kernel.

Mirror of the parts of the reference's make_ccf.py (paths under
/root/reference/py/rvspecfit/) that `fitter_ccf.fit` calls before its hot loop:
`get_ccf_config` (make_ccf.py:67-102), `get_continuum` (:105-164),
`preprocess_model` (:167-212), `preprocess_data` (:330-414) -- SURVEY.md section 8
row f3.  Two routes feed the CCF kernels of fitter_ccf.py with the
continuum-normalised spectrum and inverse variance on the CCF pixel grid:
  * `DevicePrep` -- the batched device route (csrc/ccf_prep.cu, rvs_ccf_preprocess):
    masks, gap bridging, medians, the soft-L1 continuum fit and the resampling for
    all spectra of one pixel grid in one launch, one CTA per spectrum;
  * `preprocess_data` / `preprocess_many` -- the host route, a line-by-line mirror
    of the reference (numpy / scipy); kept for single-object calls, for spectra the
    device route does not take (non-finite fluxes, too many pixels or continuum
    nodes) and as the bit-level parity check of everything around the continuum fit.
"""
import atexit
import logging
import os

import numpy as np
import scipy.interpolate
import scipy.optimize
import scipy.signal
import scipy.stats

C_CCF = 3e5      # km/s; the CCF code's rounded speed of light (fitter_ccf.py:132)


def get_ccf_config(logl0=None, logl1=None, npoints=None, splinestep=1000, maxcontpts=20):
    """CCF configuration dictionary (make_ccf.py:67-102)."""
    conf = dict(logl0=logl0, logl1=logl1, npoints=npoints, continuum=splinestep is not None,
                maxcontpts=maxcontpts)
    if splinestep is not None:
        widest = C_CCF * (np.exp((logl1 - logl0) / maxcontpts) - 1)
        conf['splinestep'] = max(splinestep, widest)
    return conf


def _continuum_model(p, nodes, lam):
    log_cont = scipy.interpolate.UnivariateSpline(nodes, p, s=0, k=2)(lam)
    return np.exp(np.clip(log_cont, -100, 100))


def get_continuum(lam0, spec0, espec0, ccfconf=None):
    """Quadratic-spline continuum in log flux with a soft-L1 loss
    (make_ccf.py:105-164)."""
    lo = lam0.min()
    dlog = np.log(1 + ccfconf['splinestep'] / C_CCF)
    nnodes = int(np.ceil(np.log(lam0.max() / lo) / dlog))
    k = np.arange(nnodes + 1)
    nodes = lo * np.exp(k[:-1] * dlog)
    edges = lo * np.exp((k - 0.5) * dlog)
    med = np.median(spec0)
    if med <= 0:
        med = abs(med) or 1
        logging.warning('The spectrum has a median that is non-positive...')
    binned = scipy.stats.binned_statistic(lam0, spec0, 'median', bins=edges).statistic
    start = np.log(np.maximum(binned, 1e-3 * med))
    start[~np.isfinite(start)] = np.log(med)
    sol = scipy.optimize.least_squares(
        lambda p: (_continuum_model(p, nodes, lam0) - spec0) / espec0, start, loss='soft_l1')
    return _continuum_model(sol['x'], nodes, lam0)


def interp_masker(lam, spec, badmask):
    """Bridge masked pixels linearly, extend the edges flat (make_ccf.py:287-327)."""
    out = np.array(spec, dtype=np.float64)
    good = np.flatnonzero(~badmask)
    bad = np.flatnonzero(badmask)
    if good.size == 0:
        logging.warning('All the pixels are masked for the ccf determination')
        out[~np.isfinite(out)] = 1
        return out
    nxt = np.searchsorted(good, bad)
    at_left, at_right = nxt == 0, nxt == good.size
    inner = ~(at_left | at_right)
    ia, ib = good[nxt[inner] - 1], good[nxt[inner]]
    la, lb, l0 = lam[ia], lam[ib], lam[bad[inner]]
    out[bad[at_left]] = spec[good[0]]
    out[bad[at_right]] = spec[good[-1]]
    out[bad[inner]] = (-(la - l0) * spec[ib] + (lb - l0) * spec[ia]) / (lb - la)
    return out


def preprocess_data(lam, spec0, espec, ccfconf=None, badmask=None, maxerr=10):
    """Continuum-normalise a spectrum and put it (and its inverse variance) on
    the CCF's log-wavelength pixels (make_ccf.py:330-414).  Returns
    (proc_spec, proc_ivar), each of length npoints."""
    grid_lam = np.exp(np.linspace(ccfconf['logl0'], ccfconf['logl1'], ccfconf['npoints']))
    err = np.array(espec, dtype=np.float64)
    flux = np.array(spec0, dtype=np.float64)
    bad = np.zeros(len(err), dtype=bool) if badmask is None else np.asarray(badmask, dtype=bool)
    smooth = scipy.signal.medfilt(flux, 11)
    typical_err = np.nanmedian(err)
    if ccfconf['continuum']:
        bad = bad | (err > maxerr * typical_err) | (smooth <= 0)
    err[bad] = 1e9 * typical_err
    flux = interp_masker(lam, flux, bad)
    cont = get_continuum(lam, flux, err, ccfconf=ccfconf) if ccfconf['continuum'] else 1
    ivar = 1. / err**2
    ivar[bad] = 0
    med = np.median(flux)
    cont = np.maximum(1e-2 * med, cont) if med > 0 else np.maximum(cont, 1)
    norm = spec0 / cont
    ivar = cont**2 * ivar
    norm[bad] = 0
    # linear interpolation onto the CCF pixels; the variance follows the weights
    left = np.searchsorted(lam, grid_lam) - 1
    inside = (left >= 0) & (left <= len(lam) - 2)
    li = left[inside]
    ri = li + 1
    wr = (grid_lam[inside] - lam[li]) / (lam[ri] - lam[li])
    wl = 1 - wr
    out_spec, out_ivar = np.zeros(len(grid_lam)), np.zeros(len(grid_lam))
    out_spec[inside] = wl * norm[li] + wr * norm[ri]
    il, ir = ivar[li], ivar[ri]
    out_ivar[inside] = il * ir / (wl**2 * ir + wr**2 * il + ((il * ir) == 0).astype(int))
    return out_spec, out_ivar


class DevicePrep:
    """Tables of one (pixel grid, CCF configuration) pair for rvs_ccf_preprocess, and the
    call itself.  The continuum model of the reference is exp(UnivariateSpline(nodes, p,
    s=0, k=2)(lam)) (make_ccf.py:118-125): linear in p, so its values at the pixels are
    Cb @ p with Cb[:, j] the spline through the j-th unit vector -- computed here ONCE per
    pixel grid with the same FITPACK routine the reference calls per residual evaluation."""

    MAX_NODES = 24

    def __init__(self, lam, ccfconf):
        from . import _cabi, _dev
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        self.npix, self.ccfconf = len(lam), dict(ccfconf)
        self.npoints = int(ccfconf['npoints'])
        self.continuum = bool(ccfconf['continuum'])
        self.ok, self.why = True, ''
        nn = 0
        Cb = np.zeros((self.npix, 1))
        bins = np.zeros(2, dtype=np.int32)
        if self.continuum:
            lo = lam.min()
            dlog = np.log(1 + ccfconf['splinestep'] / C_CCF)
            nn = int(np.ceil(np.log(lam.max() / lo) / dlog))
            k = np.arange(nn + 1)
            nodes = lo * np.exp(k[:-1] * dlog)
            edges = lo * np.exp((k - 0.5) * dlog)
            if nn < 3 or nn > self.MAX_NODES:
                self.ok, self.why = False, f'{nn} continuum nodes'
            else:
                eye = np.eye(nn)
                Cb = np.stack([scipy.interpolate.UnivariateSpline(nodes, eye[j], s=0, k=2)(lam)
                               for j in range(nn)], axis=1)
                # binned_statistic: bin k holds edges[k] <= lam < edges[k+1], the last bin
                # its right edge too
                bins = np.searchsorted(lam, edges, side='left').astype(np.int32)
                bins[-1] = np.searchsorted(lam, edges[-1], side='right')
        if not np.all(np.diff(lam) > 0):
            self.ok, self.why = False, 'wavelengths not increasing'
        if self.ok and _cabi.lib().rvs_ccf_prep_smem(self.npix, nn) > 220 * 1024:
            self.ok, self.why = False, f'{self.npix} pixels exceed the shared-memory window'
        grid_lam = np.exp(np.linspace(ccfconf['logl0'], ccfconf['logl1'], self.npoints))
        left = np.searchsorted(lam, grid_lam) - 1
        inside = (left >= 0) & (left <= self.npix - 2)
        li = np.where(inside, left, 0)
        wr = np.where(inside, (grid_lam - lam[li]) / (lam[li + 1] - lam[li]), 0.0)
        self.nn = nn
        if self.ok:
            self.d_lam = _dev.upload(lam, np.float64)
            self.d_Cb = _dev.upload(Cb, np.float64)
            self.d_bin = _dev.upload(bins, np.int32)
            self.d_left = _dev.upload(np.where(inside, left, -1), np.int32)
            self.d_wr = _dev.upload(wr, np.float64)

    def __call__(self, spec, espec, badmask=None, maxerr=10, want_cont=False):
        """spec, espec (n, npix) host arrays -- or lists of n rows, which then go row by
        row into pinned staging and up in one asynchronous copy each --, badmask (n, npix)
        bool or None -> device tensors (proc_spec, proc_ivar) of shape (n, npoints)
        [, cont, info]."""
        from . import _cabi, _dev
        torch = _dev.torch_mod()
        n = len(spec)
        if isinstance(spec, list):
            d_spec = _dev.upload_concat(spec, np.float64)[1]
            d_espec = _dev.upload_concat(espec, np.float64)[1]
            if badmask is not None:
                badmask = _dev.upload_concat(badmask, np.bool_)[1].view(torch.uint8)
        else:
            d_spec, d_espec = _dev.upload(spec, np.float64), _dev.upload(espec, np.float64)
        d_bad = None if badmask is None else badmask if torch.is_tensor(badmask) else \
            torch.from_numpy(np.ascontiguousarray(badmask, dtype=np.uint8)).to(_dev.device())
        d_ps = _dev.empty((n, self.npoints), np.float64)
        d_pi = _dev.empty((n, self.npoints), np.float64)
        d_cont = _dev.empty((n, self.npix), np.float64) if want_cont else None
        d_info = _dev.empty((n, 2), np.int32)
        rc = _cabi.lib().rvs_ccf_preprocess(
            _dev.ptr(self.d_lam), _dev.ptr(d_spec), _dev.ptr(d_espec), _dev.ptr(d_bad), n,
            self.npix, _dev.ptr(self.d_Cb), self.nn, _dev.ptr(self.d_bin), _dev.ptr(self.d_left),
            _dev.ptr(self.d_wr), self.npoints, int(self.continuum), float(maxerr), _dev.ptr(d_ps),
            _dev.ptr(d_pi), _dev.ptr(d_cont), _dev.ptr(d_info), _dev.stream())
        _cabi.check(rc, 'rvs_ccf_preprocess')
        return (d_ps, d_pi, d_cont, d_info) if want_cont else (d_ps, d_pi)


_device_preps = {}


def device_prep(lam, gridkey, ccfconf):
    """Cached DevicePrep of a pixel grid (SpecData.gridkey) and CCF configuration."""
    key = (gridkey, tuple(sorted((k, float(v)) for k, v in ccfconf.items())))
    if key not in _device_preps:
        if len(_device_preps) > 64:
            _device_preps.pop(next(iter(_device_preps)))
        _device_preps[key] = DevicePrep(lam, ccfconf)
    return _device_preps[key]


# ---- many spectra: the continuum fit is ~40 ms of scipy per arm, so a batch of survey
# size is preprocessed by a pool of host processes, one (object, arm) per task -- what the
# reference's drivers do with whole objects (desi/desi_fit.py:1475-1479, OMP_NUM_THREADS=1)
_pool = None


def _pool_init():
    try:
        import threadpoolctl
        _pool_init.limit = threadpoolctl.threadpool_limits(1)
    except ImportError:
        pass


def _preprocess_job(job):
    lam, spec, espec, badmask, ccfconf = job
    return preprocess_data(lam, spec, espec, ccfconf=ccfconf, badmask=badmask)


def _close_pool():
    global _pool
    if _pool is not None:
        _pool[1].shutdown(wait=False, cancel_futures=True)
        _pool = None


def preprocess_many(jobs, workers=None):
    """preprocess_data for a list of (lam, spec, espec, badmask, ccfconf) in a persistent
    pool of `workers` spawned processes (default: the host's cores, at most 32; serial
    below 16 jobs).  Same values as the serial loop, in job order."""
    jobs = list(jobs)
    if workers is None:
        workers = min(os.cpu_count() or 1, 32) if len(jobs) >= 16 else 1
    if workers <= 1 or len(jobs) < 2:
        return [_preprocess_job(j) for j in jobs]
    global _pool
    if _pool is None or _pool[0] != workers:
        import concurrent.futures as cf
        import multiprocessing as mp
        _close_pool()
        _pool = (workers, cf.ProcessPoolExecutor(workers, mp_context=mp.get_context('spawn'),
                                                 initializer=_pool_init))
        atexit.register(_close_pool)
    chunk = max(1, len(jobs) // (8 * workers))
    return list(_pool[1].map(_preprocess_job, jobs, chunksize=chunk))


def preprocess_model(logl, lammodel, model0, vsini=None, ccfconf=None, broaden=None):
    """Template on the CCF pixels, continuum-normalised (make_ccf.py:167-212).
    `broaden(lam, spec, vsini)` applies the rotation kernel (bank preparation is
    an offline step; pass spec_inter's device routine or a host one)."""
    m = model0
    if vsini:
        if broaden is None:
            raise ValueError('preprocess_model: vsini given without a broadening routine')
        m = broaden(lammodel, model0, vsini)
    cont = 1
    if ccfconf['continuum']:
        cont = get_continuum(lammodel, m, np.maximum(m * 1e-5, 1e-2 * np.median(m)),
                             ccfconf=ccfconf)
        cont = np.maximum(cont, 1e-2 * np.median(cont))
    ll = np.log(lammodel)
    if not (ll[0] <= logl[0] <= ll[-1]) or not (ll[0] <= logl[-1] <= ll[-1]):
        logging.warning('The required wavelength range is bigger than the template wavelengths')
    return scipy.interpolate.interp1d(ll, m / cont, bounds_error=False, fill_value=1)(logl)


def _model_job(job):
    logl, lam, spec, ccfconf = job
    return preprocess_model(logl, lam, spec, vsini=None, ccfconf=ccfconf)


def build_bank(bank, node_params, ccfconf, every=10, vsinis=(0.,), workers=None):
    """CCF template bank of a device-resident TemplateBank, in memory: every `every`-th
    row of node_params (physical units, (n, ndim)) x the vsini list, interpolated and
    rotationally broadened on the device, continuum-normalised
    and resampled to the CCF pixels (reference make_ccf.py:417-493; an offline product in
    the reference, built here because the synthetic workloads have no files).  Returns
    dict(fft, fft2, models, params, vsinis, parnames, ccfconf) for
    fitter_ccf.register_ccf_bank(name, **that)."""
    logl = np.linspace(ccfconf['logl0'], ccfconf['logl1'], ccfconf['npoints'])
    node_params = np.asarray(node_params, dtype=np.float64)
    inds = np.arange(0, len(node_params), every)
    params = node_params[inds]
    jobs, pars, vs = [], [], []
    for v in vsinis:
        spec, _ = bank.template(params, vsini=(float(v) if v else None))
        for k in range(len(inds)):
            jobs.append((logl, bank.lam, spec[k], ccfconf))
            pars.append(params[k])
            vs.append(float(v))
    if workers is None:
        workers = min(os.cpu_count() or 1, 32) if len(jobs) >= 16 else 1
    if workers > 1:
        import concurrent.futures as cf
        import multiprocessing as mp
        with cf.ProcessPoolExecutor(workers, mp_context=mp.get_context('spawn'),
                                    initializer=_pool_init) as ex:
            models = list(ex.map(_model_job, jobs, chunksize=max(1, len(jobs) // (4 * workers))))
    else:
        models = [_model_job(j) for j in jobs]
    models = np.array(models)
    # template order of the reference: node-major (make_ccf.py:468-475)
    order = np.lexsort((np.arange(len(vs)) // len(inds), np.arange(len(vs)) % len(inds)))
    models, pars, vs = models[order], np.array(pars)[order], [vs[i] for i in order]
    return dict(fft=np.fft.rfft(models, axis=1), fft2=np.fft.rfft(models**2, axis=1),
                models=models, params=pars, vsinis=vs, parnames=list(bank.parnames),
                ccfconf=dict(ccfconf))
