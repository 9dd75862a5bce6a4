"""Likelihood of a spectrum given a template, on the GPU.

Mirror of the reference's spec_fit.py (paths under
/root/reference/py/rvspecfit/): `SpecData`, `get_chisq`, `find_best`,
`get_chisq_continuum` keep their signatures, return values and error
behaviour; `LikelihoodEngine` is the batched sibling they are built on.

All arithmetic of the path -- template interpolation, vsini broadening, spline
construction, Doppler resampling, continuum normal equations, chi-square, RV
scan statistics -- runs in librvs_b200.so.  The host resolves grid vertices
(spec_inter.TemplateBank.locate), adds the additive penalty terms
(spec_fit.py:863,888-896) and raises the reference's exceptions.
"""
import ctypes
import os
import random
import threading
import time

import numpy as np
import scipy.linalg

from . import _cabi, _dev, spec_inter

SPEED_OF_LIGHT = 299792.458  # km/s, spec_fit.py:23


class ResolMatrix:
    """Holder of a resolution matrix (reference spec_fit.py:54-67): `mat` is anything
    scipy.sparse can turn into diagonals (the reference builds scipy.sparse
    dia matrices, spec_fit.py:466, desi/desi_fit.py:746)."""

    def __init__(self, mat):
        self.fd = {'mat': mat}
        self.objid = random.getrandbits(128)

    def __hash__(self):
        return self.objid

    @property
    def mat(self):
        return self.fd['mat']


def construct_resol_mat(lam, resol=None, width=None):
    """Resolution matrix of Gaussian line-spread functions from a resolving power
    R = lambda/delta lambda (sigma = lam/R/2.35) or a width in Angstrom: reference
    spec_fit.py:410-468, same arguments and result (host-side data preparation; the
    matrix is applied on the device)."""
    import scipy.sparse
    assert (resol is None or width is None)
    assert (resol is not None or width is not None)
    lam = np.asarray(lam, dtype=np.float64)
    n = len(lam)
    sigs = lam / resol / 2.35 if resol is not None else np.zeros(n) + width
    assert (np.all(np.diff(lam) > 0))
    pix = np.arange(n)
    left = np.maximum(np.searchsorted(lam, lam - 5 * sigs, 'left'), 0)
    right = np.minimum(np.searchsorted(lam, lam + 5 * sigs, 'right'), n - 1)
    half = min(n, max(np.max(right - pix), np.max(pix - left)))
    offsets = np.arange(-half, half + 1)
    # profile of pixel j sampled at its neighbours j + offsets, normalised per pixel j
    nb = pix[None, :] + offsets[:, None]
    inside = (nb >= 0) & (nb < n)
    prof = np.exp(-0.5 * ((lam[np.where(inside, nb, 0)] - lam[None, :]) / sigs[None, :])**2)
    prof = prof * inside
    prof /= prof.sum(axis=0)[None, :]
    # scipy's diagonal storage: data[k, c] = M[c - offsets[k], c], with
    # M[r, c] = prof[offsets = c - r ... as the reference lays it out] (spec_fit.py:463-465)
    cid = (pix[None, :] - offsets[:, None]) % n
    kid = np.broadcast_to((half + offsets)[:, None], cid.shape)
    return ResolMatrix(scipy.sparse.spdiags(prof[kid, cid], offsets, n, n))


def convolve_resol(spec, resol_matrix):
    """reference spec_fit.py:471-489 (host convenience; the likelihood applies the
    matrix on the device)."""
    return resol_matrix.mat @ spec


def _band_rows(mat, n):
    """Diagonals of a resolution matrix by OUTPUT pixel (rvs_obs.d_resol):
    (offsets ascending int32[nd], rows f64[nd, n]) with rows[k, p] = M[p, p + offsets[k]]."""
    import scipy.sparse
    dia = scipy.sparse.dia_matrix(mat)
    if dia.shape != (n, n):
        raise ValueError(f'resolution matrix of shape {dia.shape} for a spectrum of {n} pixels')
    offs, rows = [], []
    for k in np.argsort(dia.offsets):
        off = int(dia.offsets[k])
        p = np.arange(max(0, -off), min(n, n - off))
        p = p[p + off < dia.data.shape[1]]
        row = np.zeros(n)
        row[p] = dia.data[k, p + off]
        if off == 0 or row.any():
            offs.append(off)
            rows.append(row)
    if 0 not in offs:
        offs.append(0)
        rows.append(np.zeros(n))
        order = np.argsort(offs)
        offs, rows = [offs[i] for i in order], [rows[i] for i in order]
    return np.array(offs, dtype=np.int32), np.array(rows, dtype=np.float64)


class SpecData:
    """A single spectroscopic dataset (reference spec_fit.py:70-145)."""

    def __init__(self, name, lam, spec, espec, badmask=None, resolution=None,
                 dtype=np.float64):
        self.name = name
        self.lam = np.ascontiguousarray(lam, dtype=dtype)
        self.spec = np.ascontiguousarray(spec, dtype=dtype)
        self.espec = np.ascontiguousarray(espec, dtype=dtype)
        self.resolution = resolution
        self._band = None
        if resolution is not None:      # spectra sharing one matrix object share its band rows
            cache = getattr(resolution, '__dict__', {}).setdefault('_band_cache', {})
            if len(self.lam) not in cache:
                cache[len(self.lam)] = _band_rows(resolution.mat, len(self.lam))
            self._band = cache[len(self.lam)]
        self.spec_error_ratio = np.ascontiguousarray(spec / espec, dtype=dtype)
        if badmask is None:
            badmask = np.zeros(len(self.spec), dtype=bool)
        self.badmask = np.asarray(badmask, dtype=bool)
        for a in (self.lam, self.spec, self.espec, self.badmask):
            a.setflags(write=False)
        self.objid = random.getrandbits(128)
        # identifies the pixel grid (objects observed on the same pixels share one
        # wavelength / basis entry on the device, SpectrumBatch)
        self.gridkey = (len(self.lam), self.lam[:4].tobytes(), self.lam[-4:].tobytes(),
                        float(self.lam.sum()))

    def __hash__(self):
        return self.objid


class SpectrumBatch:
    """Ragged batch of spectra of ONE setup in device memory, with the derived
    per-pixel products and the continuum basis (spec_fit.py:103-108,148-176)."""

    def __init__(self, specdatas):
        self.n = len(specdatas)
        self.npix = np.array([len(s.lam) for s in specdatas], dtype=np.int64)
        self.max_npix = int(self.npix.max())
        self.off = np.concatenate([[0], np.cumsum(self.npix)]).astype(np.int64)
        self.lam0 = np.array([s.lam[0] for s in specdatas])
        self.lam1 = np.array([s.lam[-1] for s in specdatas])
        self.h_spec, self.d_spec = _dev.upload_concat([s.spec for s in specdatas], np.float64)
        self.h_espec, self.d_espec = _dev.upload_concat([s.espec for s in specdatas], np.float64)
        self.h_bad = np.concatenate([s.badmask for s in specdatas])
        self._lams = [s.lam for s in specdatas]
        self._gridkeys = [s.gridkey for s in specdatas]
        self.d_off = _dev.upload(self.off, np.int64)
        # objects observed on the same pixels share one wavelength grid: one copy
        # of lam / ln(lam) / continuum basis in the grid pools (rvs_obs)
        seen, gid, first = {}, np.zeros(self.n, dtype=np.int64), []
        for i, s in enumerate(specdatas):
            key = s.gridkey
            if key not in seen:
                seen[key] = len(first)
                first.append(i)
            gid[i] = seen[key]
        self.grid_of = gid
        self.grid_first = np.array(first, dtype=np.int64)
        gl = [specdatas[i].lam for i in self.grid_first]
        self.gstart = np.concatenate([[0], np.cumsum([len(_) for _ in gl])]).astype(np.int64)
        self.d_glam = _dev.upload(np.concatenate(gl), np.float64)
        self.d_gstart = _dev.upload(self.gstart, np.int64)
        self.d_goff = _dev.upload(self.gstart[:-1][gid], np.int64)
        self._prod, self._basis = {}, {}
        self._mk_lock = threading.Lock()     # derived products are made once, by one thread
        # resolution matrices: band rows by output pixel on the diagonals any member
        # has (a member without a matrix gets the identity), one object pool
        self.resol_offs = self.d_resol = self.d_resol_offs = None
        bands = [getattr(s, '_band', None) for s in specdatas]
        if any(b is not None for b in bands):
            offs = sorted({0}.union(*[set(b[0].tolist()) for b in bands if b is not None]))
            blocks = self._resol_blocks(specdatas, offs)
            self.resol_offs = np.array(offs, dtype=np.int32)
            self.h_resol, self.d_resol = _dev.upload_concat(blocks, np.float64)
            self.d_resol_offs = _dev.upload(self.resol_offs, np.int32)

    def lam_of(self, i):
        """Wavelengths of object i (host)."""
        return self._lams[i]

    def _resol_blocks(self, specdatas, offs):
        pos = {o: k for k, o in enumerate(offs)}
        blocks = []
        for s in specdatas:
            b = getattr(s, '_band', None)
            blk = np.zeros((len(offs), len(s.lam)))
            if b is None:
                blk[pos[0]] = 1.0
            else:
                if not set(b[0].tolist()) <= set(pos):
                    raise ValueError('resolution matrix with diagonals outside the batch\'s band')
                blk[[pos[o] for o in b[0].tolist()]] = b[1]
            blocks.append(blk.ravel())
        return blocks

    def reload(self, specdatas):
        """New spectra of the SAME layout (number of objects, pixel grids, resolution
        band) into the existing device buffers: fluxes and errors go through the pinned
        staging buffers this batch already owns, the derived products (and band rows) are
        recomputed in place.  Device addresses do not change, so captured CUDA graphs of
        the engine stay valid.  Stream-ordered on the current stream."""
        torch = _dev.torch_mod()
        if len(specdatas) != self.n or \
                any(len(s.lam) != n for s, n in zip(specdatas, self.npix)) or \
                any(s.gridkey != k for s, k in zip(specdatas, self._gridkeys)):
            raise ValueError('reload: the spectra do not have the layout of the batch')
        if (self.d_resol is None) != all(getattr(s, '_band', None) is None for s in specdatas):
            raise ValueError('reload: resolution matrices do not match the batch')
        torch.cuda.current_stream().synchronize()    # the staging buffers are free again
        np.concatenate([s.spec for s in specdatas], out=self.h_spec)
        np.concatenate([s.espec for s in specdatas], out=self.h_espec)
        self.d_spec.copy_(torch.from_numpy(self.h_spec), non_blocking=True)
        self.d_espec.copy_(torch.from_numpy(self.h_espec), non_blocking=True)
        _dev.IO_BYTES[0] += self.h_spec.nbytes + self.h_espec.nbytes
        self.h_bad = np.concatenate([s.badmask for s in specdatas])
        if self.d_resol is not None:
            np.concatenate(self._resol_blocks(specdatas, self.resol_offs.tolist()),
                           out=self.h_resol)
            self.d_resol.copy_(torch.from_numpy(self.h_resol), non_blocking=True)
            _dev.IO_BYTES[0] += self.h_resol.nbytes
        for key, (dn, einv, sumlog2) in self._prod.items():
            rc = _cabi.lib().rvs_obs_prepare(
                _dev.ptr(self.d_spec), _dev.ptr(self.d_espec), _dev.ptr(self.d_off), self.n, key,
                _dev.ptr(dn), _dev.ptr(einv), _dev.ptr(sumlog2), _dev.stream())
            _cabi.check(rc, 'rvs_obs_prepare')

    def products(self, sys_err=0.0):
        key = float(sys_err)
        if key in self._prod:
            return self._prod[key]
        with self._mk_lock:
            return self._products_locked(key)

    def _products_locked(self, key):
        if key not in self._prod:
            ntot = int(self.off[-1])
            dn, einv = [_dev.empty((ntot,), np.float64) for _ in range(2)]
            sumlog2 = _dev.empty((self.n,), np.float64)
            rc = _cabi.lib().rvs_obs_prepare(
                _dev.ptr(self.d_spec), _dev.ptr(self.d_espec), _dev.ptr(self.d_off), self.n, key,
                _dev.ptr(dn), _dev.ptr(einv), _dev.ptr(sumlog2), _dev.stream())
            _cabi.check(rc, 'rvs_obs_prepare')
            self._prod[key] = (dn, einv, sumlog2)
        return self._prod[key]

    def basis(self, npoly, rbf):
        """(loglam, P) of the grid pools; P is pixel-major [pixel][npp]."""
        key = (int(npoly), bool(rbf))
        if key in self._basis:
            return self._basis[key]
        with self._mk_lock:
            return self._basis_locked(key, npoly, rbf)

    def _basis_locked(self, key, npoly, rbf):
        if key not in self._basis:
            ntot = int(self.gstart[-1])
            npp = (npoly + 1) // 2 * 2
            P = _dev.empty((ntot, npp), np.float64)
            loglam = _dev.empty((ntot,), np.float64)
            rc = _cabi.lib().rvs_basis_build(_dev.ptr(self.d_glam), _dev.ptr(self.d_gstart),
                                             len(self.grid_first), ntot, int(npoly),
                                             int(bool(rbf)), npp, _dev.ptr(loglam), _dev.ptr(P),
                                             _dev.stream())
            _cabi.check(rc, 'rvs_basis_build')
            self._basis[key] = (loglam, P, npp)
        return self._basis[key]

    @property
    def tn_rows(self):
        """Rows of the T/sigma workspace per item (rvs_chisq_fused: two with
        resolution matrices)."""
        return 1 if self.d_resol is None else 2

    def obs(self, npoly, rbf, sys_err=0.0, resol=True):
        """struct rvs_obs (the tensors it points to stay alive in the caches)."""
        dn, einv, sumlog2 = self.products(sys_err)
        loglam, P, npp = self.basis(npoly, rbf)
        o = _cabi.Obs()
        o.d_lam, o.d_loglam, o.d_P = self.d_glam.data_ptr(), loglam.data_ptr(), P.data_ptr()
        o.d_goff, o.d_dn, o.d_einv = self.d_goff.data_ptr(), dn.data_ptr(), einv.data_ptr()
        o.d_sumlog2, o.d_off = sumlog2.data_ptr(), self.d_off.data_ptr()
        o.npoly, o.npp, o.nobj = int(npoly), npp, self.n
        o.shared_grid = int(len(self.grid_first) == 1)
        if resol and self.d_resol is not None:
            o.d_resol, o.d_resol_offs = self.d_resol.data_ptr(), self.d_resol_offs.data_ptr()
            o.nresol = len(self.resol_offs)
            o.resol_halfwidth = int(np.abs(self.resol_offs).max())
        return o


# kernels launched through CUDA-graph replays by all engines of this process
# (rvs_launch_count only sees direct launches)
GRAPH_LAUNCHES = [0]
# priority of the evaluation slots' streams (rvs_stream_create: 0 least, 1 halfway, 2 greatest)
SLOT_STREAM_PRIORITY = int(os.environ.get('RVS_SLOT_PRIO', '0'))
LAST_DRIVE_ROUNDS = [0]     # rounds of the last native Nelder-Mead stage (tests)


class PendingEval:
    """Handle of a LikelihoodEngine.submit() call."""

    def ready(self):
        """True when result() would not wait for the device."""
        return self.slot is None or self.slot['event'].query()

    def result(self):
        eng, obj, vels, params, vs, outside_penalty, espec_systematic, raise_errors = self.args
        if self.slot is None:
            return eng._evaluate_general(obj, vels, params, vs, outside_penalty,
                                         espec_systematic, False, raise_errors)
        total, redo = eng._collect_fast(self.slot, obj, vels, outside_penalty)
        self.slot = None
        if redo.any():
            r = np.nonzero(redo)[0]
            eng.n_eval -= len(r)
            total[r] = eng._evaluate_general(obj[r], vels[r], params[r],
                                             None if vs is None else vs[r], outside_penalty,
                                             espec_systematic, False, raise_errors)
        return total


class PendingFit:
    """Handle of a LikelihoodEngine.submit_fit() call."""

    def ready(self):
        return self.slot['event'].query()

    def result(self):
        """(values (K,), redo (K,) bool): objective values (1e30 behind the walls) and
        the items the fused path could not settle."""
        sl, eng, K = self.slot, self.eng, self.K
        sl['event'].synchronize()
        eng._account(sl)
        narm = len(eng.setups)
        n = _cabi.lib().rvs_fit_collect(
            ctypes.byref(self.lay), K, sl['Kp'], _dev.hptr(self.obj32),
            ctypes.c_void_p(sl['h_in'].data_ptr()), ctypes.c_void_p(sl['h_chi'].data_ptr()),
            ctypes.c_void_p(sl['h_flags'].data_ptr()), int(bool(sl.get('shared_locate'))), 1,
            _dev.hptr(sl['f_prior']), _dev.hptr(sl['f_pen']), _dev.hptr(sl['f_wall']),
            _dev.hptr(sl['f_out']), _dev.hptr(sl['f_redo']))
        out = sl['f_out'][:K].copy()
        redo = sl['f_redo'][:K].astype(bool) if n else None
        sl['busy'] = False
        if n < 0:
            _cabi.check(int(n), 'rvs_fit_collect')
        return out, redo


def _overlap_ok(t0, t1, s0, s1, vmin, vmax):
    """spec_fit.py:786-794, vectorised; True where the template covers the data."""
    ok = np.ones(np.broadcast(s0, vmin).shape, dtype=bool)
    for v in (vmin, vmax):
        corr = np.sqrt((1 + v / SPEED_OF_LIGHT) / (1 - v / SPEED_OF_LIGHT))
        ok &= ~((t0 * corr > s0) | (t1 * corr < s1))
    return ok


class LikelihoodEngine:
    """Batched -2 log L over many objects.

    objects: list of objects, each a list of SpecData (one per arm; arms may
    differ between objects).  Every evaluation call takes `obj` (K,) indices
    into that list, so that one object can be evaluated at many trial points in
    one launch.
    """

    def __init__(self, objects, config, options=None, fused=True):
        options = options or {}
        self.config = config
        self.npoly = options.get('npoly') or 5
        self.rbf = options.get('rbf_continuum', True)
        self.fused = fused
        if self.npoly > _cabi.MAX_NPOLY:
            raise ValueError(f'npoly={self.npoly} > {_cabi.MAX_NPOLY} is not supported')
        self.objects = [[o] if isinstance(o, SpecData) else list(o) for o in objects]
        self.nobj = len(self.objects)
        self.setups = []
        for o in self.objects:
            for sd in o:
                if sd.name not in self.setups:
                    self.setups.append(sd.name)
        self.badchi = np.array([10 * sum(len(sd.lam) for sd in o) for o in self.objects],
                               dtype=np.float64)
        self.arms = {}
        for name in self.setups:
            members, index = [], np.full(self.nobj, -1, dtype=np.int64)
            for i, o in enumerate(self.objects):
                for sd in o:
                    if sd.name == name:
                        index[i] = len(members)
                        members.append(sd)
            bank = spec_inter.getInterpolator(name, config).bank
            self.arms[name] = dict(batch=SpectrumBatch(members), index=index, bank=bank)
        self.parnames = self.arms[self.setups[0]]['bank'].parnames
        self.n_eval = 0
        self.timer = None
        self.use_graphs = not os.environ.get('RVS_NO_GRAPHS')
        self.graph_kernel_launches = 0      # kernels launched through graph replays
        # fast path (see _evaluate_fast): per-arm object index on the host, and
        # whether the template covers each object over [min_vel, max_vel]
        # (spec_fit.py:786-794 evaluated once instead of per call)
        self._oix = np.stack([self.arms[n]['index'] for n in self.setups]).astype(np.int32)
        cov = np.ones((len(self.setups), self.nobj), dtype=bool)
        for a, n in enumerate(self.setups):
            arm = self.arms[n]
            has = arm['index'] >= 0
            ix = arm['index'][has]
            cov[a, has] = _overlap_ok(arm['bank'].lam[0], arm['bank'].lam[-1],
                                      arm['batch'].lam0[ix], arm['batch'].lam1[ix],
                                      config['min_vel'], config['max_vel'])
        self._cover0 = cov.all(axis=0)
        self._fast_banks = all(
            self.arms[n]['bank'].kind == 'regulargrid' and self.arms[n]['bank'].gridmap is not None
            and self.arms[n]['bank'].knots.ratio_dev < 1e-8 for n in self.setups)
        self._buf = {}
        self._obs_cache = {}
        self._data_epoch, self._data_event = 0, None    # see _launch
        self._lock = threading.RLock()      # slots, scratch buffers, captures, counters
        self._general_lock = threading.RLock()   # the general path's shared workspaces
        self._launch_skew = 0
        # largest vsini whose rotation kernel fits the fused path on every arm
        # (tapcap = ceil(v / c / lnstep + 1) + 1 <= MAX_FUSED_TAPS)
        self._fused_vmax = min(
            (_cabi.MAX_FUSED_TAPS - 2) * 299792.458 * self.arms[n]['bank'].knots.lnstep
            for n in self.setups) * (1 - 1e-12)

    def reload(self, objects):
        """Replace the spectra by new ones of the same layout (same arms per object, same
        pixel grids): SpectrumBatch.reload for every arm.  The engine, its device buffers
        and its captured CUDA graphs stay -- the streaming pattern of a survey driver that
        pushes exposure after exposure of one instrument through one engine."""
        objects = [[o] if isinstance(o, SpecData) else list(o) for o in objects]
        if len(objects) != self.nobj or any(
                [sd.name for sd in a] != [sd.name for sd in b]
                for a, b in zip(objects, self.objects)):
            raise ValueError('reload: the objects do not have the layout of the engine')
        if any(sl['busy'] for sl in getattr(self, '_slots', [])):
            raise RuntimeError('reload with evaluations in flight: collect their results first')
        for name in self.setups:
            arm = self.arms[name]
            arm['batch'].reload([sd for o in objects for sd in o if sd.name == name])
        self.objects = objects
        self._data_epoch += 1
        self._data_event = None

    @staticmethod
    def _fusable(bank, vs):
        """The fused kernel needs exactly (log-)uniform knots and a rotation
        kernel of at most MAX_FUSED_TAPS one-sided taps (rvs_b200.h)."""
        if not bank.knots.ratio_dev < 1e-8:
            return False
        vmax = 0.0 if vs is None else float(np.max(vs, initial=0.0))
        return bank.tapcap(vmax) <= _cabi.MAX_FUSED_TAPS

    def _tap_bound(self, vsini):
        """Upper bound of vsini that sizes the tap buffers of a fused call, in coarse
        steps (powers of two from 16 km/s) so that consecutive evaluations share one
        launch configuration.  The rounded value is clamped to the largest one whose
        rotation kernel still fits the fused path wherever the actual maximum does, so
        that submit()'s gate and rvs_chisq_fused agree."""
        vmax = 0.0 if vsini is None else float(vsini.max(initial=0.0))
        if not vmax > 0:
            return 0.0
        rounded = float(max(16.0, 2.0 ** np.ceil(np.log2(vmax))))
        return rounded if rounded <= self._fused_vmax else vmax

    def drain(self):
        """Wait for every evaluation in flight and free its slot (used when a driver
        stops early, e.g. on an exception, with submitted work it will not collect)."""
        for sl in getattr(self, '_slots', []):
            if sl['busy']:
                sl['event'].synchronize()
                sl['busy'] = False

    def _workspace(self, n):
        if getattr(self, '_ws', None) is None or self._ws.numel() < n:
            self._ws = _dev.empty((int(n * 1.25) + 1024,), np.float64)
        return self._ws

    # ------------------------------------------------------------------ core
    def _arm_eval(self, arm, sel, obj, vels, params, vsini, sys_err, want_model,
                  fast_interp=False):
        """chi-square of one arm for the items `sel` (indices into the call's
        item list).  vels (K, nv).  Returns chisq (k, nv), status (k, nv),
        outside (k,), tstatus (k,), extras."""
        bank, batch = arm['bank'], arm['batch']
        L = _cabi.lib()
        k = len(sel)
        nv = vels.shape[1]
        ids, w, outside = bank.locate(params[sel])
        oix = arm['index'][obj[sel]].astype(np.int32)
        obs = batch.obs(self.npoly, self.rbf, sys_err)
        d_oix = _dev.upload(oix, np.int32)
        d_vels = _dev.upload(vels[sel], np.float64)
        d_chi = _dev.empty((k, nv), np.float64)
        d_st = _dev.empty((k, nv), np.int32)
        vs = None if vsini is None else np.ascontiguousarray(vsini[sel], dtype=np.float64)
        extras = None
        if self.fused and nv == 1 and not want_model and not fast_interp and \
                self._fusable(bank, vs):
            d_ids = _dev.upload(ids, np.int32)
            d_w = _dev.upload(w, np.float64)
            d_vs = None if vs is None else _dev.upload(vs, np.float64)
            vmax = 0.0 if vs is None else float(np.max(vs, initial=0.0))
            stride = int(batch.npix.max())
            d_tn = self._workspace(k * stride * batch.tn_rows)
            nwork = L.rvs_fused_workspace(k, bank.tapcap(vmax), bank.npix_t)
            if getattr(self, '_work', None) is None or self._work.numel() < nwork:
                self._work = _dev.empty((int(nwork * 1.25) + 64,), np.float64)
            d_work = self._work
            t0 = self.timer.start() if self.timer else None
            rc = L.rvs_chisq_fused(_dev.ptr(bank.grid), bank.grid_f64, bank.ld,
                                   ctypes.byref(bank.knots), _dev.ptr(d_ids), _dev.ptr(d_w),
                                   bank.nvert, _dev.ptr(d_vs), vmax, int(bank.log_spec),
                                   ctypes.byref(obs), _dev.ptr(d_oix), _dev.ptr(d_vels), k,
                                   _dev.ptr(d_tn), stride, _dev.ptr(d_work), _dev.ptr(d_chi),
                                   _dev.ptr(d_st),
                                   ctypes.byref(bank.box) if bank.box is not None else None,
                                   _dev.stream())
            _cabi.check(rc, 'rvs_chisq_fused')
            if t0 is not None:
                self.timer.stop('fused', t0, k)
            chi = _dev.download(d_chi)
            st = _dev.download(d_st)
            redo = np.nonzero(st[:, 0] & _cabi.ST_LIMIT)[0]
            if len(redo):      # window did not fit the fused path: general path
                self.fused = False
                try:
                    c2, s2, _, t2, _ = self._arm_eval(arm, sel[redo], obj, vels, params, vsini,
                                                      sys_err, False)
                finally:
                    self.fused = True
                chi[redo], st[redo] = c2, s2 | t2[:, None]
            tstatus = st[:, 0] & (_cabi.ST_TEMPLATE_BAD | _cabi.ST_TAPS)
            return chi, st, outside, tstatus, None
        else:
            t0 = self.timer.start() if self.timer else None
            yz, d_tst = bank.build(ids, w, vs)
            if t0 is not None:
                self.timer.stop('build', t0, k)
            d_tix = _dev.upload(np.arange(k), np.int32)
            d_co = d_raw = d_mod = d_moff = None
            if want_model:
                npx = batch.npix[oix]
                moff = np.concatenate([[0], np.cumsum(npx)]).astype(np.int64)
                d_moff = _dev.upload(moff[:-1], np.int64)
                d_co = _dev.empty((k, self.npoly), np.float64)
                d_raw = _dev.empty((int(moff[-1]),), np.float64)
                d_mod = _dev.empty((int(moff[-1]),), np.float64)
            t0 = self.timer.start() if self.timer else None
            rc = L.rvs_chisq_scan(_dev.ptr(yz), bank.npix_t, _dev.ptr(d_tix),
                                  ctypes.byref(bank.knots), ctypes.byref(obs), _dev.ptr(d_oix),
                                  _dev.ptr(d_vels), nv, k, _dev.ptr(d_chi), _dev.ptr(d_st),
                                  _dev.ptr(d_co), _dev.ptr(d_raw), _dev.ptr(d_mod),
                                  _dev.ptr(d_moff), int(bool(fast_interp)), _dev.stream())
            _cabi.check(rc, 'rvs_chisq_scan')
            if t0 is not None:
                self.timer.stop('scan', t0, k * nv)
            st = _dev.download(d_st)
            tstatus = _dev.download(d_tst)
            if want_model:
                extras = dict(coeffs=_dev.download(d_co), raw=_dev.download(d_raw),
                              model=_dev.download(d_mod), moff=moff, oix=oix)
        return _dev.download(d_chi), st, outside, tstatus, extras

    def _epoch(self, sl):
        """What the validity of a slot's captured graphs hangs on: the engine's shared
        buffers and the slot's own (a slot that grows does not invalidate the others)."""
        return (getattr(self, '_graph_epoch', 0), sl.get('epoch', 0))

    def _scratch(self, name, shape, dtype, reserve=0, sl=None):
        """Device buffer reused between calls (grown by 25 % when too small; `reserve`:
        elements to allocate at least, e.g. the slot's capacity)."""
        n = int(np.prod(shape))
        t = self._buf.get(name)
        if t is None or t.numel() < n:
            if t is not None:
                _dev.torch_mod().cuda.synchronize()   # side streams may still read the old one
            t = _dev.empty((max(int(n * 1.25) + 64, int(reserve)),), dtype)
            self._buf[name] = t
            if sl is not None:          # captured pointers are stale
                sl['epoch'] = sl.get('epoch', 0) + 1
            else:
                self._graph_epoch = getattr(self, '_graph_epoch', 0) + 1
        return t[:n].view(*shape)

    NSLOT = 16     # evaluations that may be in flight at once (submit without result)

    def _slot(self, K, narm, nd):
        """Pinned host staging + device input buffers of one in-flight evaluation."""
        torch = _dev.torch_mod()
        if not hasattr(self, '_slots'):
            self._slots, self._slot_ix = [dict(K=0, busy=False, ix=i) for i in range(self.NSLOT)], -1
        free = [x for x in self._slots if not x['busy']]
        if not free:
            return None
        sl = free[0]    # lowest free slot: its captured graphs are the warmest
        # every slot is sized for the largest call any slot has seen, so that which
        # slot a call lands in (it varies with the timing of concurrent drivers) never
        # decides whether buffers grow and captured graphs are dropped
        self._kcap = max(getattr(self, '_kcap', 0), K)
        if sl['K'] < K:
            sl['epoch'] = sl.get('epoch', 0) + 1
            cap = max(int(K * 1.25) + 16, self._slot_cap())
            pin = dict(pin_memory=True)
            sl.update(K=cap,
                      h_in=torch.empty(((2 + nd) * cap,), dtype=torch.float64, **pin),
                      h_oix=torch.empty((narm * cap,), dtype=torch.int32, **pin),
                      h_chi=torch.empty((2 * narm * cap,), dtype=torch.float64, **pin),
                      h_flags=torch.empty((2 * narm * cap,), dtype=torch.int32, **pin),
                      d_in=_dev.empty(((2 + nd) * cap,), np.float64),
                      d_oix=_dev.empty((narm * cap,), np.int32),
                      event=torch.cuda.Event())
        return sl

    def _slot_cap(self):
        return int(self._kcap * 1.25) + 16

    def _acquire(self, K):
        """Round the item count up to a launch configuration and take a free slot
        (pinned staging, device inputs, streams) for it: (slot, Kp).  With every slot
        taken by other threads' evaluations the call waits for one to be collected; a
        single-threaded driver that never collects gets the error instead."""
        L = _cabi.lib()
        narm = len(self.setups)
        bank0 = self.arms[self.setups[0]]['bank']
        nd = bank0.ndim
        if not hasattr(self, '_same_maps'):
            self._same_maps = all(self.arms[n]['bank'].log_ids == bank0.log_ids
                                  for n in self.setups)
        use_graph = self.use_graphs and self._same_maps and \
            not getattr(self, 'serial_arms', False) and not L.rvs_profile_active()
        # The item count is rounded up (absent items: arm index -1 everywhere, skipped by
        # every kernel), so that an optimiser whose active set shrinks call by call keeps
        # hitting the same captured launch configurations.
        Kp = K
        if use_graph:
            Kp = int(L.rvs_fit_round_items(K))
        torch = _dev.torch_mod()
        me = threading.get_ident()
        while True:
            with self._lock:
                sl = self._slot(Kp, narm, nd)
                if sl is not None:
                    sl['busy'], sl['owner'] = True, me
                    break
                if all(x.get('owner') == me for x in self._slots):
                    raise RuntimeError(f'more than {self.NSLOT} evaluations in flight: call '
                                       'result() on the oldest PendingEval first')
            time.sleep(2e-5)
        sl['Kp'], sl['use_graph'] = Kp, use_graph
        # Every in-flight evaluation has its own streams and scratch, so that the
        # low-occupancy tail of one (continuum solves of the last arm) runs under the
        # template kernels of the next instead of in front of them.
        if 'stream' not in sl:
            # streams of the slot's own (torch.cuda.Stream() hands out the handles of a
            # small pool round-robin: slots would share streams)
            def own_stream():
                h = L.rvs_stream_create(SLOT_STREAM_PRIORITY)
                if not h:
                    raise _cabi.RvsError('rvs_stream_create failed')
                return torch.cuda.ExternalStream(h)
            sl['stream'] = own_stream()
            sl['ready'] = torch.cuda.Event()
            sl['arm_streams'] = [own_stream() for _ in self.setups]
            sl['arm_events'] = [torch.cuda.Event() for _ in self.setups]
            sl['fork_event'] = torch.cuda.Event()
        return sl, Kp

    def _submit_fast(self, obj, vels, params, vsini, sys_errs, vmax):
        """Optimiser-phase evaluation of K items at one velocity each with no
        host work per item and no host synchronisation: one asynchronous upload
        of (vel, vsini, mapped parameters) from pinned memory, then per arm vertex
        location, rotation taps, the fused template/resampling kernel and the
        continuum solve, all stream-ordered; one asynchronous download of the
        per-arm chi-squares and flags into pinned memory, and an event.  Items that
        need anything else (off-grid or missing-corner points, template not
        finite, normal matrix not PD, velocity outside [min_vel, max_vel],
        template not covering the data) come back flagged and are re-evaluated by
        the general path when the result is collected.  Returns the slot."""
        K = len(obj)
        narm = len(self.setups)
        bank0 = self.arms[self.setups[0]]['bank']
        nd = bank0.ndim
        sl, Kp = self._acquire(K)
        # inputs of this evaluation -> pinned staging (uploaded by the enqueued work)
        host_in = sl['h_in'][:(2 + nd) * Kp].view(2 + nd, Kp).numpy()
        q = spec_inter.map_params(params, bank0.log_ids).T
        host_in[0, :K] = vels
        host_in[1, :K] = 0.0 if vsini is None else vsini
        host_in[2:, :K] = q
        h_oix = sl['h_oix'][:narm * Kp].view(narm, Kp).numpy()
        h_oix[:, :K] = self._oix[:, obj]
        if Kp > K:
            host_in[:2, K:] = 0.0
            host_in[2:, K:] = q[:, :1]
            h_oix[:, K:] = -1
        self._launch(sl, K, Kp, vmax, sys_errs, params)
        return sl

    def _obs_all(self, sys_errs):
        """struct rvs_obs of every arm for the given systematic errors; the derived
        products behind them are made once, by one thread (on its current stream: slot
        streams wait for them through the data event, see _launch)."""
        obs_all = self._obs_cache.get(sys_errs)
        if obs_all is None:
            with self._lock:
                obs_all = self._obs_cache.get(sys_errs)
                if obs_all is None:
                    obs_all = [self.arms[name]['batch'].obs(self.npoly, self.rbf, sys_errs[a])
                               for a, name in enumerate(self.setups)]
                    _dev.torch_mod().cuda.current_stream().synchronize()
                    self._obs_cache[sys_errs] = obs_all
                    self._data_epoch += 1
                    self._data_event = None
        return obs_all

    def _wait_data(self, sl):
        torch = _dev.torch_mod()
        if sl.get('data_epoch') != self._data_epoch:
            with self._lock:
                if self._data_event is None:
                    self._data_event = torch.cuda.Event()
                    self._data_event.record(torch.cuda.current_stream())
                sl['stream'].wait_event(self._data_event)
                sl['data_epoch'] = self._data_epoch

    def _launch(self, sl, K, Kp, vmax, sys_errs, params):
        """Start the device work of the evaluation whose inputs are in the slot's pinned
        buffers; records the slot's event behind it."""
        L = _cabi.lib()
        torch = _dev.torch_mod()
        narm = len(self.setups)
        nd = self.arms[self.setups[0]]['bank'].ndim
        use_graph = sl['use_graph']
        # per-arm data products are made (once) on the caller's stream
        okey = tuple(sys_errs)
        obs_all = self._obs_all(okey)
        # The slot's stream waits for the spectra and their derived products (uploaded /
        # computed on the caller's stream by the constructor, reload() or the first obs()
        # of a systematic error) once per such change -- not for whatever else the
        # caller's stream carries (another lock-step set's RV scan, say)
        self._wait_data(sl)
        # The ~20 launches, copies and stream fork/joins of one evaluation are captured
        # into a CUDA graph the second time a configuration (item count, tap bound,
        # systematic error) is seen and replayed from then on: one launch per evaluation
        # instead of a host-bound launch sequence.
        key = (Kp, vmax, okey)
        epoch = self._epoch(sl)
        if sl.get('graph_epoch') != epoch:
            sl['graphs'], sl['seen'], sl['graph_epoch'] = {}, {}, epoch
        g = sl['graphs'].get(key) if use_graph else None
        with torch.cuda.stream(sl['stream']):
            t0 = self.timer.start() if self.timer else None
            if g is not None:
                g[0].replay()
                nk = g[1]
            else:
                # direct launches and graph captures: one thread at a time (scratch buffers
                # may grow, which moves the graph epoch)
                with self._lock:
                    nk, done = 0, False
                    if use_graph and sl['seen'].get(key, 0) >= 1 and len(sl['graphs']) < 192:
                        # second sighting: every scratch buffer of this configuration exists
                        l0 = L.rvs_launch_count()
                        gr = torch.cuda.CUDAGraph()
                        gr.capture_begin(capture_error_mode='thread_local')
                        try:
                            self._enqueue_fast(sl, obs_all, params, vmax, Kp, narm, nd)
                        finally:
                            gr.capture_end()
                        nk = L.rvs_launch_count() - l0
                        self._launch_skew -= nk                 # captured, not run
                        if self._epoch(sl) != epoch:   # a buffer moved during the capture
                            sl['graphs'], sl['seen'] = {}, {}
                            sl['graph_epoch'] = self._epoch(sl)
                            nk = 0
                        else:
                            sl['graphs'][key] = (gr, nk)
                            gr.replay()
                            done = True
                    if not done:
                        self._enqueue_fast(sl, obs_all, params, vmax, Kp, narm, nd)
                        nk = 0      # counted by the library
                    sl['seen'][key] = sl['seen'].get(key, 0) + 1
            if t0 is not None:
                self.timer.stop('fused_eval', t0, K)
            sl['event'].record()
        sl['graph_kernels'] = nk

    def _account(self, sl):
        """Counters of a finished evaluation (kernels launched through graph replays,
        bytes moved)."""
        narm, Kp = len(self.setups), sl['Kp']
        nd = self.arms[self.setups[0]]['bank'].ndim
        nk = sl.get('graph_kernels', 0)
        with self._lock:
            self.graph_kernel_launches += nk
            GRAPH_LAUNCHES[0] += nk + self._launch_skew
            self._launch_skew = 0
            _dev.IO_BYTES[0] += (2 + nd) * Kp * 8 + narm * Kp * 4
            _dev.IO_BYTES[1] += 2 * narm * Kp * (8 + 4)

    def _enqueue_fast(self, sl, obs_all, params, vmax, K, narm, nd):
        """Enqueue (or capture) the device work of one evaluation on the current
        stream: uploads, vertex location, the arms on their own streams, downloads."""
        L = _cabi.lib()
        bank0 = self.arms[self.setups[0]]['bank']
        six = sl['ix']
        d_in = sl['d_in'][:(2 + nd) * K].view(2 + nd, K)
        d_oix = sl['d_oix'][:narm * K].view(narm, K)
        d_in.copy_(sl['h_in'][:(2 + nd) * K].view(2 + nd, K), non_blocking=True)
        d_oix.copy_(sl['h_oix'][:narm * K].view(narm, K), non_blocking=True)
        cap = sl['K']       # scratch is sized for the slot's capacity: it grows when the slot does
        d_chi = self._scratch(f'chi_{six}', (2, narm, K), np.float64, 2 * narm * cap, sl=sl)    # chi-square | off-grid measure
        d_flags = self._scratch(f'flags_{six}', (2, narm, K), np.int32, 2 * narm * cap, sl=sl)
        nvert = bank0.nvert
        torch = _dev.torch_mod()
        main = torch.cuda.current_stream()
        # the arms are independent: each runs its kernel sequence on its own stream,
        # so the small kernels of one arm (vertex location, preparation, solves) fill
        # the SMs left idle by the tails and launch gaps of the others
        banks = [self.arms[name]['bank'] for name in self.setups]
        sigs = [getattr(b, 'locate_signature', None) for b in banks]
        shared = narm > 1 and sigs[0] is not None and all(s_ == sigs[0] for s_ in sigs)
        sl['shared_locate'] = shared
        if shared:      # one vertex location for all arms, before the streams fork
            d_ids0 = self._scratch(f'ids0_{six}', (K, nvert), np.int32, cap * nvert, sl=sl)
            d_w0 = self._scratch(f'w0_{six}', (K, nvert), np.float64, cap * nvert, sl=sl)
            rc = L.rvs_locate_grid(ctypes.byref(bank0.gridmap), _dev.ptr(d_in[2:]), K, K,
                                   _dev.ptr(d_ids0), _dev.ptr(d_w0), _dev.ptr(d_flags[1, 0]),
                                   _dev.ptr(d_chi[1, 0]), ctypes.c_void_p(main.cuda_stream))
            _cabi.check(rc, 'rvs_locate_grid')
        if shared and narm <= _cabi.MAX_ARMS and not getattr(self, 'serial_arms', False) \
                and not os.environ.get('RVS_NO_MERGE'):
            # all arms in ONE launch of every kernel of the call (rvs_chisq_fused_multi)
            arms = (_cabi.FusedArm * narm)()
            keep = []
            for a, name in enumerate(self.setups):
                arm = self.arms[name]
                bank, batch = arm['bank'], arm['batch']
                stride = batch.max_npix
                d_tn = self._scratch(f'tn{a}_{six}', (K * stride * batch.tn_rows,), np.float64,
                                     cap * stride * batch.tn_rows, sl=sl)
                d_work = self._scratch(
                    f'work{a}_{six}',
                    (L.rvs_fused_workspace(K, bank.tapcap(vmax), bank.npix_t),), np.float64,
                    L.rvs_fused_workspace(cap, bank.tapcap(vmax), bank.npix_t), sl=sl)
                fa = arms[a]
                fa.d_grid, fa.grid_f64, fa.log_spec = bank.grid.data_ptr(), bank.grid_f64, \
                    int(bank.log_spec)
                fa.ld = bank.ld
                fa.knots = ctypes.pointer(bank.knots)
                fa.obs = ctypes.pointer(obs_all[a])
                fa.d_oix, fa.d_tn, fa.tn_stride = d_oix[a].data_ptr(), d_tn.data_ptr(), stride
                fa.d_work = d_work.data_ptr()
                fa.d_chisq, fa.d_status = d_chi[0, a].data_ptr(), d_flags[0, a].data_ptr()
                if bank.box is not None:
                    fa.box = ctypes.pointer(bank.box)
                keep.append((d_tn, d_work))
            rc = L.rvs_chisq_fused_multi(arms, narm, _dev.ptr(d_ids0), _dev.ptr(d_w0), nvert,
                                         _dev.ptr(d_in[1]) if vmax > 0 else None, vmax,
                                         _dev.ptr(d_in[0]), K,
                                         ctypes.c_void_p(main.cuda_stream))
            _cabi.check(rc, 'rvs_chisq_fused_multi')
        else:
            self._enqueue_arms(sl, obs_all, params, vmax, K, narm, d_in, d_oix, d_chi, d_flags,
                               shared, d_ids0 if shared else None, d_w0 if shared else None)
        sl['h_chi'][:2 * narm * K].view(2, narm, K).copy_(d_chi, non_blocking=True)
        sl['h_flags'][:2 * narm * K].view(2, narm, K).copy_(d_flags, non_blocking=True)

    def _enqueue_arms(self, sl, obs_all, params, vmax, K, narm, d_in, d_oix, d_chi, d_flags,
                      shared, d_ids0, d_w0):
        """The arms of a call on their own streams, one rvs_chisq_fused each (arms whose
        banks do not share a node table, or that cannot share launches)."""
        L = _cabi.lib()
        torch = _dev.torch_mod()
        main = torch.cuda.current_stream()
        bank0 = self.arms[self.setups[0]]['bank']
        six = sl['ix']
        nvert = bank0.nvert
        sl['fork_event'].record(main)
        for a, name in enumerate(self.setups):
            arm = self.arms[name]
            bank, batch = arm['bank'], arm['batch']
            obs = obs_all[a]
            # serial_arms (stage profiling): every arm on the caller's stream
            st = main if getattr(self, 'serial_arms', False) else sl['arm_streams'][a]
            st.wait_event(sl['fork_event'])
            stream = ctypes.c_void_p(st.cuda_stream)
            if shared:
                d_ids, d_w = d_ids0, d_w0
            else:
                d_ids = self._scratch(f'ids{a}_{six}', (K, nvert), np.int32, sl=sl)
                d_w = self._scratch(f'w{a}_{six}', (K, nvert), np.float64, sl=sl)
                q = d_in[2:]
                if bank.log_ids != bank0.log_ids:
                    with torch.cuda.stream(st):
                        q = _dev.upload(spec_inter.map_params(params, bank.log_ids).T, np.float64)
                rc = L.rvs_locate_grid(ctypes.byref(bank.gridmap), _dev.ptr(q), K, K,
                                       _dev.ptr(d_ids), _dev.ptr(d_w), _dev.ptr(d_flags[1, a]),
                                       _dev.ptr(d_chi[1, a]), stream)
                _cabi.check(rc, 'rvs_locate_grid')
            stride = batch.max_npix
            d_tn = self._scratch(f'tn{a}_{six}', (K * stride * batch.tn_rows,), np.float64, sl=sl)
            d_work = self._scratch(f'work{a}_{six}',
                                   (L.rvs_fused_workspace(K, bank.tapcap(vmax), bank.npix_t),),
                                   np.float64, sl=sl)
            rc = L.rvs_chisq_fused(_dev.ptr(bank.grid), bank.grid_f64, bank.ld,
                                   ctypes.byref(bank.knots), _dev.ptr(d_ids), _dev.ptr(d_w),
                                   bank.nvert, _dev.ptr(d_in[1]) if vmax > 0 else None, vmax,
                                   int(bank.log_spec), ctypes.byref(obs), _dev.ptr(d_oix[a]),
                                   _dev.ptr(d_in[0]), K, _dev.ptr(d_tn), stride, _dev.ptr(d_work),
                                   _dev.ptr(d_chi[0, a]), _dev.ptr(d_flags[0, a]),
                                   ctypes.byref(bank.box) if bank.box is not None else None,
                                   stream)
            _cabi.check(rc, 'rvs_chisq_fused')
            sl['arm_events'][a].record(st)
        for a in range(narm):
            main.wait_event(sl['arm_events'][a])

    def _collect_fast(self, sl, obj, vels, outside_penalty=True):
        """Wait for a submitted evaluation: (total (K,), redo (K,) bool)."""
        K, narm, Kp = len(obj), len(self.setups), sl['Kp']
        sl['event'].synchronize()
        self._account(sl)
        both = sl['h_chi'][:2 * narm * Kp].view(2, narm, Kp).numpy()[:, :, :K]
        chi, outside = both[0], both[1]
        flags = sl['h_flags'][:2 * narm * Kp].view(2, narm, Kp).numpy()[:, :, :K]
        if sl.get('shared_locate'):     # located once: arm 0's rows stand for every arm
            outside = np.broadcast_to(outside[0], outside.shape)
            both = np.stack([chi, outside])
            flags[1, 1:] = flags[1, 0]
        redo = (flags != 0).any(axis=(0, 1)) | ~np.isfinite(both).all(axis=(0, 1))
        redo |= ~self._cover0[obj] | (vels < self.config['min_vel']) | \
            (vels > self.config['max_vel'])
        if outside_penalty and outside.any():
            # off-grid points were resolved on the device (nearest node); their
            # penalty is added once per arm the object HAS, as get_chisq does
            # (spec_fit.py:879,895-896: one term per SpecData of the object)
            present = self._oix[:, obj] >= 0
            chi = np.where(present, outside * self.badchi[obj][None, :], 0.0) + chi
        total = np.add.reduce(chi, axis=0)
        sl['busy'] = False
        return total, redo

    def scan(self, obj, V, nv, params, vsini=None, quadratic=True):
        """find_best for K objects with one template each (spec_fit.py:1018-1092), without
        taking the chi-square matrices through the host: V (K, nvmax) velocities (row k uses
        its first nv[k]), params (K, ndim), vsini (K,) or None.  Vertex location, template
        build and the RV-scan GEMM per arm, the sum over arms and the statistics kernel all
        run on the device; one (K, 8) array comes back (rvs_scan_stats: best_chi, best_vel,
        vel_err, skewness, kurtosis, i_vel, i_par, flags).  Returns (stats, redo): redo (K,)
        marks the objects this route could not settle (missing arms, off-grid template not
        usable, template not covering the data, a non-finite or flagged chi-square) -- their
        rows of `stats` are undefined and the caller takes the general path for them.
        Returns None when the engine's banks do not allow the route at all."""
        if not self._fast_banks:
            return None
        L = _cabi.lib()
        torch = _dev.torch_mod()
        obj = np.asarray(obj, dtype=np.int64)
        V = np.ascontiguousarray(V, dtype=np.float64)
        K, nvmax = V.shape
        nv = np.asarray(nv, dtype=np.int32)
        params = np.array(params, dtype=np.float64, ndmin=2)
        narm = len(self.setups)
        banks = [self.arms[n]['bank'] for n in self.setups]
        bank0 = banks[0]
        sig0 = getattr(bank0, 'locate_signature', None)
        if sig0 is None or any(getattr(b, 'locate_signature', None) != sig0 for b in banks):
            return None
        min_vel, max_vel = self.config['min_vel'], self.config['max_vel']
        redo = ~self._cover0[obj] | (self._oix[:, obj] < 0).any(axis=0)
        with np.errstate(invalid='ignore'):
            redo |= (V.min(axis=1) < min_vel) | (V.max(axis=1) > max_vel)
        d_V = _dev.upload(V, np.float64)
        d_q = _dev.upload(spec_inter.map_params(params, bank0.log_ids).T, np.float64)
        d_vs = None if vsini is None else _dev.upload(np.asarray(vsini, dtype=np.float64),
                                                       np.float64)
        d_ids = _dev.empty((K, bank0.nvert), np.int32)
        d_w = _dev.empty((K, bank0.nvert), np.float64)
        d_lflag = _dev.empty((K,), np.int32)
        d_out_dist = _dev.empty((K,), np.float64)
        stream = _dev.stream()
        rc = L.rvs_locate_grid(ctypes.byref(bank0.gridmap), _dev.ptr(d_q), K, K, _dev.ptr(d_ids),
                               _dev.ptr(d_w), _dev.ptr(d_lflag), _dev.ptr(d_out_dist), stream)
        _cabi.check(rc, 'rvs_locate_grid')
        d_bad = _dev.upload(self.badchi[obj], np.float64)
        d_pen = d_out_dist * d_bad                 # off-grid penalty of one arm (spec_fit.py:895-896)
        d_tix = torch.arange(K, dtype=torch.int32, device=d_V.device)
        d_tot = torch.zeros((K, nvmax), dtype=torch.float64, device=d_V.device)
        d_flag = d_lflag != 0
        with self._lock:
            self.n_eval += int(nv.sum()) * 1
        obs_all = self._obs_all((0.0,) * narm)
        ragged = bool((nv != nvmax).any())
        d_nv = _dev.upload(nv, np.int32)
        for a, name in enumerate(self.setups):
            arm = self.arms[name]
            bank, batch = arm['bank'], arm['batch']
            obs = obs_all[a]
            d_oix = _dev.upload(np.maximum(self._oix[a, obj], 0), np.int32)
            yz = _dev.empty((K, bank.npix_t, 2), np.float64)
            d_tst = _dev.empty((K,), np.int32)
            t0 = self.timer.start() if self.timer else None
            rc = L.rvs_template_build(
                _dev.ptr(bank.grid), bank.grid_f64, bank.ld, ctypes.byref(bank.knots),
                _dev.ptr(d_ids), _dev.ptr(d_w), bank.nvert, _dev.ptr(d_vs), int(bank.log_spec),
                K, _dev.ptr(yz), bank.npix_t, _dev.ptr(d_tst), stream)
            _cabi.check(rc, 'rvs_template_build')
            if t0 is not None:
                self.timer.stop('build', t0, K)
            d_chi = _dev.empty((K, nvmax), np.float64)
            d_st = _dev.empty((K, nvmax), np.int32)
            t0 = self.timer.start() if self.timer else None
            # ragged refinement grids: the trials past an object's own count are skipped
            rc = L.rvs_chisq_scan_ragged(_dev.ptr(yz), bank.npix_t, _dev.ptr(d_tix),
                                         ctypes.byref(bank.knots), ctypes.byref(obs),
                                         _dev.ptr(d_oix), _dev.ptr(d_V), nvmax,
                                         _dev.ptr(d_nv) if ragged else None, K, _dev.ptr(d_chi),
                                         _dev.ptr(d_st), None, None, None, None, 0, stream)
            _cabi.check(rc, 'rvs_chisq_scan_ragged')
            if t0 is not None:
                self.timer.stop('scan', t0, int(nv.sum()))
            # total += penalty + chi-square of the arm, in the order the general path adds
            d_tot += d_pen[:, None] + d_chi
            d_flag |= (d_tst != 0) | (d_st != 0).any(dim=1)
        d_stats = _dev.empty((K, 8), np.float64)
        rc = L.rvs_scan_stats_ragged(_dev.ptr(d_V), _dev.ptr(d_tot), K, 1, nvmax, _dev.ptr(d_nv),
                                     int(quadratic), _dev.ptr(d_stats), None, stream)
        _cabi.check(rc, 'rvs_scan_stats_ragged')
        stats = _dev.download(d_stats)
        redo |= _dev.download(d_flag.to(torch.uint8)).astype(bool)
        redo |= ~np.isfinite(stats[:, 0])
        return stats, redo

    # ---- native round loop of a Nelder-Mead stage (csrc/drive_host.cpp) ----
    def drive_open(self, lay, objmap, nfit, cap):
        """Take an evaluation slot for a whole optimiser stage whose requests have at
        most `cap` points of `nfit` coordinates, problems numbered like objmap (int32
        engine objects).  Returns the stage handle for drive_run / drive_close, or None
        when the packed fast path does not apply."""
        if not (self.fused and self._fast_banks and self.use_graphs):
            return None
        L = _cabi.lib()
        sl, Kp = self._acquire(cap)
        if not self._same_maps or not sl['use_graph']:
            sl['busy'] = False
            return None
        if sl.get('fit_cap', 0) < Kp:
            c = sl['K']
            sl.update(fit_cap=c, f_prior=np.empty(c), f_pen=np.empty(c),
                      f_wall=np.empty(c, dtype=np.uint8), f_out=np.empty(c),
                      f_redo=np.empty(c, dtype=np.uint8))
        self._wait_data(sl)
        h = L.rvs_drive_create(int(cap), int(nfit))
        if not h:
            sl['busy'] = False
            raise _cabi.RvsError('rvs_drive_create failed')
        io = _cabi.Drive()
        io.stream = sl['stream'].cuda_stream
        io.h_in, io.h_oix = sl['h_in'].data_ptr(), sl['h_oix'].data_ptr()
        io.h_chi, io.h_flags = sl['h_chi'].data_ptr(), sl['h_flags'].data_ptr()
        io.f_prior, io.f_pen = sl['f_prior'].ctypes.data, sl['f_pen'].ctypes.data
        io.f_wall, io.f_out = sl['f_wall'].ctypes.data, sl['f_out'].ctypes.data
        io.f_redo = sl['f_redo'].ctypes.data
        io.cap = min(sl['K'], sl['fit_cap'])
        objmap = np.ascontiguousarray(objmap, dtype=np.int32)
        io.objmap, io.nprob = objmap.ctypes.data, len(objmap)
        io.fused_vmax = self._fused_vmax
        io.state = _cabi.DRIVE_IDLE
        st = dict(sl=sl, h=ctypes.c_void_p(h), io=io, lay=lay, objmap=objmap, nfit=int(nfit),
                  table=None, table_key=None, done=[0, 0, 0, 0, 0, 0], trec=None)
        timer = self.timer
        if timer is not None and getattr(timer, 'epoch', None) is not None:
            st['trec'] = np.zeros((1 << 16, 3))
            io.epoch_event = timer.epoch.cuda_event
            io.t_rec, io.t_cap, io.t_n = st['trec'].ctypes.data, len(st['trec']), 0
        return st

    def _drive_table(self, st):
        """The slot's captured graphs (systematic error 0) as the arrays rvs_nm_drive reads."""
        sl, io = st['sl'], st['io']
        narm = len(self.setups)
        graphs = sl.get('graphs', {}) if sl.get('graph_epoch') == self._epoch(sl) else {}
        key = (sl.get('graph_epoch'), len(graphs))
        if st['table_key'] == key:
            return
        zero = (0.0,) * narm
        ent = [(k[0], k[1], g) for k, g in graphs.items() if k[2] == zero]
        kp = np.array([e[0] for e in ent], dtype=np.int64)
        vm = np.array([e[1] for e in ent], dtype=np.float64)
        ex = (ctypes.c_void_p * max(1, len(ent)))(*[e[2][0].raw_cuda_graph_exec() for e in ent])
        nk = np.array([e[2][1] for e in ent], dtype=np.int32)
        st['table'], st['table_key'] = (kp, vm, ex, nk), key
        io.ngraph = len(ent)
        io.g_kp, io.g_vmax, io.g_nk = kp.ctypes.data, vm.ctypes.data, nk.ctypes.data
        io.g_exec = ctypes.cast(ex, ctypes.c_void_p).value

    def _drive_flush(self, st):
        """Counters of the rounds run so far -> the engine's."""
        io, done = st['io'], st['done']
        now = [io.items, io.graph_kernels, io.h2d_bytes, io.d2h_bytes, io.rounds, io.t_n]
        with self._lock:
            self.n_eval += now[0] - done[0]
            self.graph_kernel_launches += now[1] - done[1]
            GRAPH_LAUNCHES[0] += now[1] - done[1]
            _dev.IO_BYTES[0] += now[2] - done[2]
            _dev.IO_BYTES[1] += now[3] - done[3]
        if st['trec'] is not None and now[5] > done[5] and self.timer is not None:
            self.timer.add_native('fused_eval', st['trec'][done[5]:now[5]].copy())
            if now[5] >= len(st['trec']) - 1:
                io.t_n = 0
                now[5] = 0
        st['done'] = now
        LAST_DRIVE_ROUNDS[0] = int(io.rounds)

    def drive_run(self, st, nm, speculate_below, stop_stopped, redo_values, py_values,
                  kind='nm'):
        """Run rounds of the stepper `nm` (rvs_nm_* handle, or rvs_bfgs_* with kind='bfgs') on the held slot
        until every problem has stopped (returns _cabi.DRIVE_DONE) or `stop_stopped`
        problems have (DRIVE_PEEL).  redo_values(obj32, X) -> objective values through the
        general path for the items the fused path could not settle; py_values(obj32, X)
        -> objective values of a whole request that does not fit the fused path."""
        L = _cabi.lib()
        sl, io, lay = st['sl'], st['io'], st['lay']
        io.speculate_below, io.stop_stopped = int(speculate_below), int(stop_stopped)
        narm = len(self.setups)
        okey = (0.0,) * narm
        try:
            while True:
                self._drive_table(st)
                io.shared_locate = int(bool(sl.get('shared_locate')))
                rc = (L.rvs_bfgs_drive if kind == 'bfgs' else L.rvs_nm_drive)(
                    nm, st['h'], ctypes.byref(lay), ctypes.byref(io))
                if rc in (_cabi.DRIVE_DONE, _cabi.DRIVE_PEEL):
                    return rc
                if rc == _cabi.DRIVE_LAUNCH:
                    sl['Kp'] = int(io.Kp)
                    self._launch(sl, int(io.K), int(io.Kp), float(io.vmax), okey, None)
                    nk = sl.get('graph_kernels', 0)
                    with self._lock:
                        self.graph_kernel_launches += nk
                        GRAPH_LAUNCHES[0] += nk + self._launch_skew
                        self._launch_skew = 0
                    io.state = _cabi.DRIVE_LAUNCHED
                elif rc in (_cabi.DRIVE_REDO, _cabi.DRIVE_PYEVAL):
                    K, N = int(io.K), st['nfit']
                    pX, pobj = ctypes.c_void_p(), ctypes.c_void_p()
                    L.rvs_drive_request(st['h'], None, ctypes.byref(pX), ctypes.byref(pobj))
                    X = np.ctypeslib.as_array(ctypes.cast(pX, ctypes.POINTER(ctypes.c_double)),
                                              (K, N))
                    obj = np.ctypeslib.as_array(ctypes.cast(pobj, ctypes.POINTER(ctypes.c_int32)),
                                                (K,))
                    if rc == _cabi.DRIVE_REDO:
                        r = np.nonzero(sl['f_redo'][:K])[0]
                        sl['f_out'][r] = redo_values(obj[r].copy(), X[r].copy())
                        with self._lock:
                            self.n_eval -= len(r)
                    else:
                        sl['f_out'][:K] = py_values(obj.copy(), X.copy())
                        with self._lock:
                            self.n_eval -= K        # counted by the route that evaluated them
                        io.state = _cabi.DRIVE_COLLECTED
                else:
                    _cabi.check(rc, 'rvs_nm_drive')
                    raise _cabi.RvsError(f'rvs_nm_drive returned {rc}')
        finally:
            self._drive_flush(st)

    def drive_close(self, st):
        _cabi.lib().rvs_drive_destroy(st['h'])
        st['sl']['busy'] = False

    def submit_fit(self, lay, obj32, X, logvals):
        """Optimiser-phase evaluation of the batched fit's objective for K (object,
        fitted vector) pairs: rvs_fit_pack turns the vectors straight into the call's
        pinned upload buffers (no per-item numpy), the captured graph of the call is
        launched, and PendingFit.result() reduces the downloads with rvs_fit_collect.
        `lay`: _cabi.FitLayout of the objective (batch_fit.BatchObjective.layout()).
        Returns None when the fused path does not apply (the caller then uses submit)."""
        if not (self.fused and self._fast_banks):
            return None
        L = _cabi.lib()
        K = len(obj32)
        narm = len(self.setups)
        nd = lay.nspec
        sl, Kp = self._acquire(K)
        if not self._same_maps:
            sl['busy'] = False
            return None
        if sl.get('fit_cap', 0) < Kp:
            cap = sl['K']
            sl.update(fit_cap=cap, f_prior=np.empty(cap), f_pen=np.empty(cap),
                      f_wall=np.empty(cap, dtype=np.uint8), f_out=np.empty(cap),
                      f_redo=np.empty(cap, dtype=np.uint8))
        vsmax = ctypes.c_double(0.0)
        rc = L.rvs_fit_pack(ctypes.byref(lay), K, Kp, _dev.hptr(obj32), _dev.hptr(X),
                            None if logvals is None else _dev.hptr(logvals),
                            ctypes.c_void_p(sl['h_in'].data_ptr()),
                            ctypes.c_void_p(sl['h_oix'].data_ptr()), _dev.hptr(sl['f_prior']),
                            _dev.hptr(sl['f_pen']), _dev.hptr(sl['f_wall']), ctypes.byref(vsmax))
        if rc:
            sl['busy'] = False
            _cabi.check(rc, 'rvs_fit_pack')
        vmax = vsmax.value
        if vmax > 0:
            rounded = float(max(16.0, 2.0 ** np.ceil(np.log2(vmax))))
            vmax = rounded if rounded <= self._fused_vmax else vmax
        if vmax > self._fused_vmax:
            sl['busy'] = False
            return None
        with self._lock:
            self.n_eval += K
        self._launch(sl, K, Kp, vmax, (0.0,) * narm, None)
        pend = PendingFit()
        pend.eng, pend.slot, pend.lay, pend.obj32, pend.K = self, sl, lay, obj32, K
        return pend

    def submit(self, obj, vels, params, vsini=None, outside_penalty=True,
               espec_systematic=None, raise_errors=False):
        """Asynchronous `evaluate` for one velocity per item: enqueues the work
        and returns a PendingEval whose result() gives the chi-squares.  Up to
        NSLOT evaluations may be in flight, so that a driver stepping several
        groups of objects keeps the GPU busy while it digests results."""
        obj = np.asarray(obj, dtype=np.int64)
        params = np.array(params, dtype=np.float64, ndmin=2)
        vels = np.asarray(vels, dtype=np.float64)
        vs = None if vsini is None else np.asarray(vsini, dtype=np.float64)
        fast = vels.ndim == 1 and self.fused and self._fast_banks and len(obj) > 0
        if fast:
            vmax = self._tap_bound(vs)
            fast = vmax <= self._fused_vmax
        pend = PendingEval()
        pend.args = (self, obj, vels, params, vs, outside_penalty, espec_systematic, raise_errors)
        pend.slot = None
        if fast:
            if isinstance(espec_systematic, dict):
                sys_errs = [float(espec_systematic[n]) for n in self.setups]
            else:
                sys_errs = [float(espec_systematic or 0.0)] * len(self.setups)
            self.n_eval += len(obj)
            pend.slot = self._submit_fast(obj, vels, params, vs, sys_errs, vmax)
        return pend

    def evaluate(self, obj, vels, params, vsini=None, outside_penalty=True,
                 espec_systematic=None, want_model=False, raise_errors=False,
                 fast_interp=False):
        """-2 log L for K items.  obj (K,), vels (K,) or (K, nv), params (K, ndim),
        vsini (K,) or None.  Returns chisq with the shape of vels (plus an info
        dict when want_model).  Semantics of reference get_chisq
        (spec_fit.py:860-989) per item."""
        obj = np.asarray(obj, dtype=np.int64)
        params = np.array(params, dtype=np.float64, ndmin=2)
        vels = np.asarray(vels, dtype=np.float64)
        flat = vels.ndim == 1
        if flat and not want_model and not fast_interp:
            return self.submit(obj, vels, params, vsini, outside_penalty, espec_systematic,
                               raise_errors).result()
        return self._evaluate_general(obj, vels, params, vsini, outside_penalty,
                                      espec_systematic, want_model, raise_errors, fast_interp)

    def _evaluate_general(self, obj, vels, params, vsini, outside_penalty, espec_systematic,
                          want_model, raise_errors, fast_interp=False):
        """The general path, one thread at a time (it shares workspaces)."""
        with self._general_lock:
            return self._evaluate_general_impl(obj, vels, params, vsini, outside_penalty,
                                               espec_systematic, want_model, raise_errors,
                                               fast_interp)

    def _evaluate_general_impl(self, obj, vels, params, vsini, outside_penalty,
                               espec_systematic, want_model, raise_errors, fast_interp=False):
        """The general path: host vertex location (any interpolator kind, off-grid
        nearest node), any number of velocities per item, model output, SVD
        rescue, the reference's exceptions."""
        flat = vels.ndim == 1
        v2 = vels[:, None] if flat else vels
        K, nv = v2.shape
        self.n_eval += K * nv
        total = np.zeros((K, nv))
        info = dict(arms={}) if want_model else None
        min_vel, max_vel = self.config['min_vel'], self.config['max_vel']
        for name in self.setups:
            arm = self.arms[name]
            sel = np.nonzero(arm['index'][obj] >= 0)[0]
            if len(sel) == 0:
                continue
            if isinstance(espec_systematic, dict):
                sys_err = float(espec_systematic[name])
            else:
                sys_err = float(espec_systematic or 0.0)
            chi, st, outside, tstatus, extras = self._arm_eval(
                arm, sel, obj, v2, params, vsini, sys_err, want_model, fast_interp)
            bad = self.badchi[obj[sel]]
            # template unusable: outside not finite, or off-grid and not finite/huge
            tbad = ~np.isfinite(outside) | ((outside > 0) & ((tstatus & 1) != 0))
            pen = np.where(tbad, 1000 * bad, (outside * bad) if outside_penalty else 0.0)
            pen = np.where(np.isfinite(pen), pen, 1000 * bad)
            batch = arm['batch']
            oix = arm['index'][obj[sel]]
            vlo = np.minimum(min_vel, v2[sel].min(axis=1))
            vhi = np.maximum(max_vel, v2[sel].max(axis=1))
            cover = _overlap_ok(arm['bank'].lam[0], arm['bank'].lam[-1], batch.lam0[oix],
                                batch.lam1[oix], vlo, vhi) | tbad
            if raise_errors and not cover.all():
                i = int(np.nonzero(~cover)[0][0])
                raise RuntimeError(
                    f"The template library ({arm['bank'].lam[0]},{arm['bank'].lam[-1]})  "
                    f"doesn't cover this wavelength range ({batch.lam0[oix[i]]},"
                    f"{batch.lam1[oix[i]]}) with velocities {vlo[i]} {vhi[i]}")
            notfin = ~np.isfinite(chi) & ~tbad[:, None]
            if notfin.any():
                chi = self._svd_rescue(arm, sel, obj, v2, params, vsini, sys_err, chi, notfin)
                notfin = ~np.isfinite(chi) & ~tbad[:, None]
                # spec_fit.py:963-974: tolerated only off-grid with a finite template
                skip = notfin & (outside > 0)[:, None]
                if raise_errors and (notfin & ~skip).any():
                    raise RuntimeError('The log(likelihood) value is not finite'
                                       f'when processing spectral configuration {name}')
                chi = np.where(skip, 0.0, chi)
            chi = np.where(tbad[:, None], 0.0, chi)
            chi = np.where(cover[:, None], chi, np.nan)
            total[sel] += pen[:, None] + chi
            if want_model:
                info['arms'][name] = dict(sel=sel, extras=extras, tbad=tbad, outside=outside)
        out = total[:, 0] if flat else total
        return (out, info) if want_model else out

    def _svd_rescue(self, arm, sel, obj, v2, params, vsini, sys_err, chi, notfin):
        """SVD route of the reference (spec_fit.py:255-303, 337-354) for the rare
        items whose normal matrix was not positive definite on the device.  The
        resampled template comes from the GPU; only the npoly x npoly
        factorisation is redone on the host."""
        chi = chi.copy()
        batch = arm['batch']
        for r, c in zip(*np.nonzero(notfin)):
            one = np.array([sel[r]])
            _, _, _, _, ex = self._arm_eval(arm, one, obj, v2[:, c:c + 1], params, vsini,
                                            sys_err, True)
            o = int(ex['oix'][0])
            sl = slice(batch.off[o], batch.off[o + 1])
            es = batch.h_espec[sl]
            if sys_err:
                es = np.sqrt(sys_err**2 + es**2)
            polys = get_poly_basis(batch.lam_of(o), self.npoly, self.rbf)
            chi[r, c] = _chisq0_svd(batch.h_spec[sl], ex['raw'], polys, es)[0]
        return chi


def get_poly_basis(lam, npoly, rbf=True):
    """Host copy of the continuum basis (spec_fit.py:148-176); used only by the
    SVD rescue and by callers that want the basis itself."""
    t = (lam - lam[0]) / (lam[-1] - lam[0]) * 2 - 1
    P = np.zeros((npoly, len(lam)))
    if not rbf:
        for i in range(npoly):
            c = np.zeros(npoly)
            c[i] = 1
            P[i] = np.polynomial.Chebyshev(c)(t)
        return P
    for i in range(min(3, npoly)):
        P[i] = t**i
    nr = npoly - 3
    if nr > 0:
        cen = np.linspace(-1, 1, nr, True)
        P[3:] = np.exp(-0.5 * (t[None, :] - cen[:, None])**2 / (1. / nr)**2)
    return P


def _chisq0_svd(spec, templ, polys, espec):
    """spec_fit.py:255-303."""
    D = spec / espec
    G = (templ / espec)[None, :] * polys
    v = G @ D[:, None]
    u, s, vt = scipy.linalg.svd(G @ G.T, check_finite=False)
    a = vt.T @ ((1. / s)[:, None] * u.T) @ v
    chisq = np.sum(np.log(s)) + 2 * np.log(espec).sum() + np.linalg.norm(D - a.T @ G)**2
    return chisq, a.flatten()


# ------------------------------------------------------ reference-shaped API
_engine_cache = {}


def _engine_for(specdata, config, options):
    cfgkey = tuple(sorted((k, str(v)) for k, v in config.items()))
    key = (tuple(sd.objid for sd in specdata), cfgkey, (options or {}).get('npoly') or 5,
           (options or {}).get('rbf_continuum', True))
    if key not in _engine_cache:
        if len(_engine_cache) > 16:
            _engine_cache.pop(next(iter(_engine_cache)))
        _engine_cache[key] = LikelihoodEngine([list(specdata)], config, options)
    return _engine_cache[key]


_resol_views = {}


def _with_resol_params(specdata, resol_params):
    """get_chisq's resol_params dictionary (spec_fit.py:922-929): the same data with the
    matrix of its setup attached (cached, so that repeated calls reuse one engine)."""
    out = []
    for sd in specdata:
        if sd.resolution is not None:
            raise ValueError('You are not allowed to set resol_param together with'
                             'the resolution of each SpecData')
        rm = resol_params[sd.name]
        key = (sd.objid, getattr(rm, 'objid', id(rm)))
        if key not in _resol_views:
            if len(_resol_views) > 64:
                _resol_views.pop(next(iter(_resol_views)))
            _resol_views[key] = SpecData(sd.name, sd.lam, sd.spec, sd.espec, badmask=sd.badmask,
                                         resolution=rm)
        out.append(_resol_views[key])
    return out


def param_dict_to_tuple(paramDict, setup, config):
    """spec_fit.py:730-736."""
    it = spec_inter.getInterpolator(setup, config)
    return tuple(paramDict[_] for _ in it.parnames)


def get_chisq(specdata, vel, atm_params, rot_params=None, resol_params=None, options=None,
              config=None, cache=None, full_output=False, fast_interp=False,
              espec_systematic=None, outside_penalty=True):
    """-2 log L of the dataset at a velocity, atmospheric and rotation
    parameters: reference spec_fit.py:797-989, same arguments and returns.
    `cache` is accepted and ignored (the spline never leaves the device)."""
    if isinstance(specdata, SpecData):
        specdata = [specdata]
    if resol_params is not None:
        specdata = _with_resol_params(specdata, resol_params)
    eng = _engine_for(specdata, config, options or {})
    vs = None if rot_params is None else np.array([rot_params[0]], dtype=np.float64)
    par = np.array([tuple(atm_params)], dtype=np.float64)
    if not full_output:
        return float(eng.evaluate([0], np.array([float(vel)]), par, vs,
                                  outside_penalty=outside_penalty,
                                  espec_systematic=espec_systematic, raise_errors=True,
                                  fast_interp=fast_interp)[0])
    chi, info = eng.evaluate([0], np.array([float(vel)]), par, vs,
                             outside_penalty=outside_penalty,
                             espec_systematic=espec_systematic, want_model=True,
                             raise_errors=True, fast_interp=fast_interp)
    ret = dict(chisq=float(chi[0]), logl=-0.5 * float(chi[0]), chisq_array=[],
               red_chisq_array=[], npix_array=[], models=[], raw_models=[])
    for sd in specdata:
        a = info['arms'][sd.name]
        if a['tbad'][0]:
            ret['chisq_array'].append(np.nan)
            ret['red_chisq_array'].append(np.nan)
            ret['models'].append(np.zeros(len(sd.lam)) + np.nan)
            continue
        ex = a['extras']
        model, raw = ex['model'][:len(sd.lam)], ex['raw'][:len(sd.lam)]
        dev = (model - sd.spec) / sd.espec
        good = ~sd.badmask
        c = float(np.sum(dev[good]**2))
        ret['models'].append(model)
        ret['raw_models'].append(raw)
        ret['chisq_array'].append(c)
        ret['npix_array'].append(int(good.sum()))
        ret['red_chisq_array'].append(c / good.sum())
    return ret


def get_chisq_continuum(specdata, options=None):
    """Continuum-only fit (spec_fit.py:739-783): the chi-square kernel with a
    unit template."""
    options = options or {}
    npoly = options.get('npoly') or 5
    rbf = options.get('rbf_continuum', True)
    if isinstance(specdata, SpecData):
        specdata = [specdata]
    L = _cabi.lib()
    ca, ra = np.zeros(len(specdata)), np.zeros(len(specdata))
    for i, sd in enumerate(specdata):
        batch = SpectrumBatch([sd])
        # a SpecData with a resolution matrix gets it applied to the unit template, as the
        # reference does (spec_fit.py:765-767): the rows of the matrix do not sum to exactly 1
        obs = batch.obs(npoly, rbf)
        # unit template on a 4-knot linear grid covering the data: y=1, z=0
        x = np.linspace(sd.lam[0] * 0.5, sd.lam[-1] * 2, 4)
        h, hinv, cp, winv = np.zeros(3), np.zeros(3), np.zeros(2), np.zeros(2)
        L.rvs_knot_tables(_dev.hptr(x), 4, _dev.hptr(h), _dev.hptr(hinv), _dev.hptr(cp),
                          _dev.hptr(winv))
        kn = _cabi.Knots()
        L.rvs_knot_info(_dev.hptr(x), 4, 0, ctypes.byref(kn))
        tabs = [_dev.upload(_, np.float64) for _ in (x, h, hinv, cp, winv)]
        kn.d_lam_t, kn.d_h, kn.d_hinv, kn.d_cp, kn.d_winv = [t.data_ptr() for t in tabs]
        yz = _dev.upload(np.tile([1.0, 0.0], (1, 4, 1)), np.float64)
        d_z32 = _dev.zeros((1,), np.int32)
        d_vel = _dev.zeros((1, 1), np.float64)
        d_chi, d_st = _dev.empty((1, 1), np.float64), _dev.empty((1, 1), np.int32)
        d_co = _dev.empty((1, npoly), np.float64)
        d_raw, d_mod = _dev.empty((len(sd.lam),), np.float64), _dev.empty((len(sd.lam),), np.float64)
        d_moff = _dev.zeros((1,), np.int64)
        rc = L.rvs_chisq_scan(_dev.ptr(yz), 4, _dev.ptr(d_z32), ctypes.byref(kn),
                              ctypes.byref(obs), _dev.ptr(d_z32), _dev.ptr(d_vel), 1, 1,
                              _dev.ptr(d_chi), _dev.ptr(d_st), _dev.ptr(d_co), _dev.ptr(d_raw),
                              _dev.ptr(d_mod), _dev.ptr(d_moff), 0, _dev.stream())
        _cabi.check(rc, 'rvs_chisq_scan')
        dev = (_dev.download(d_mod) - sd.spec) / sd.espec
        good = ~sd.badmask
        ca[i] = np.sum(dev[good]**2)
        ra[i] = ca[i] / good.sum()
    return dict(chisq_array=ca, redchisq_array=ra)


def scan_stats(vel_grid, chisq, quadratic=True, nv=None, want_probs=True):
    """find_best tail (spec_fit.py:1072-1092) on the device.  vel_grid (S, nv),
    chisq (S, npar, nv); `nv` (S,) optional: scan s uses its first nv[s] velocities
    only (ragged refinement grids).  Returns (out (S, 8), probs (S, nv) or None)."""
    vel_grid = np.asarray(vel_grid, dtype=np.float64)
    S, npar, nvmax = chisq.shape
    d_v, d_c = _dev.upload(vel_grid, np.float64), _dev.upload(chisq, np.float64)
    d_out = _dev.empty((S, 8), np.float64)
    d_pr = _dev.empty((S, nvmax), np.float64) if want_probs else None
    if nv is None:
        rc = _cabi.lib().rvs_scan_stats(_dev.ptr(d_v), _dev.ptr(d_c), S, npar, nvmax,
                                        int(quadratic), _dev.ptr(d_out), _dev.ptr(d_pr),
                                        _dev.stream())
    else:
        d_nv = _dev.upload(np.asarray(nv), np.int32)
        rc = _cabi.lib().rvs_scan_stats_ragged(_dev.ptr(d_v), _dev.ptr(d_c), S, npar, nvmax,
                                               _dev.ptr(d_nv), int(quadratic), _dev.ptr(d_out),
                                               _dev.ptr(d_pr), _dev.stream())
    _cabi.check(rc, 'rvs_scan_stats')
    return _dev.download(d_out), (_dev.download(d_pr) if want_probs else None)


def find_best(specdata, vel_grid, params_list, rot_params=None, resol_params=None,
              options=None, config=None, quadratic=True):
    """Best template and velocity on a grid: reference spec_fit.py:1018-1092,
    same arguments and returned keys."""
    if isinstance(specdata, SpecData):
        specdata = [specdata]
    if resol_params is not None:
        specdata = _with_resol_params(specdata, resol_params)
    eng = _engine_for(specdata, config, options or {})
    vel_grid = np.asarray(vel_grid, dtype=np.float64)
    npar, nv = len(params_list), len(vel_grid)
    par = np.array([tuple(p) for p in params_list], dtype=np.float64)
    vs = None if rot_params is None else np.full(npar, rot_params[0], dtype=np.float64)
    chisq = eng.evaluate(np.zeros(npar, dtype=np.int64), np.tile(vel_grid, (npar, 1)), par, vs,
                         raise_errors=True)                      # (npar, nv)
    out, probs = scan_stats(vel_grid[None, :], chisq[None, :, :], quadratic)
    o = out[0]
    i1, i2 = int(o[5]), int(o[6])
    if quadratic and 0 < i1 < nv - 1:
        assert vel_grid[i1 - 1] < o[1] < vel_grid[i1 + 1]      # spec_fit.py:1014
    return dict(best_chi=float(o[0]), best_vel=float(o[1]), vel_err=float(o[2]),
                best_param=params_list[i2], kurtosis=float(o[4]), skewness=float(o[3]),
                probs=probs[0])
