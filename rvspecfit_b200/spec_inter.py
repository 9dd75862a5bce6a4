"""Template banks resident in HBM and the host-side vertex location.

Mirror of the reference's spec_inter.py (paths under
/root/reference/py/rvspecfit/): `getInterpolator`, `SpecInterpolator.eval /
outsideFlag`, `getSpecParams` keep their names and meaning.  The interpolation
itself (corner-weighted sum, exp) runs in rvs_template_build /
rvs_chisq_fused; the host only resolves which grid rows and weights an
evaluation needs -- 16 for the polylinear grid (spec_inter.py:134-194), ndim+1
for the Delaunay product (spec_inter.py:35-59) -- and the off-grid measure
(spec_inter.py:62-92).  Both are vectorised over a batch of parameter vectors.
"""
import ctypes
import itertools
import os

import numpy as np
import scipy.spatial

from . import _cabi, _dev


def map_params(p, log_ids):
    """LogParamMapper.forward (read_grid.py:127-145) for an (K, ndim) array."""
    q = np.array(p, dtype=np.float64, ndmin=2)
    with np.errstate(invalid='ignore', divide='ignore'):
        for i in log_ids:
            q[:, i] = np.log10(q[:, i])
    return q


class TemplateBank:
    """One spectral setup's template grid in device memory.

    dats: (Nnode, Npix) float32 (or float64) log-flux rows, as the reference's
    `interpdat_<setup>.npy` holds them.  kind 'regulargrid' needs uvecs, idgrid,
    vecs; kind 'triangulation' needs triang (scipy Delaunay), extraflags.
    """

    def __init__(self, name, lam, dats, parnames, kind='regulargrid', uvecs=None,
                 idgrid=None, vecs=None, triang=None, extraflags=None, log_ids=(0,),
                 log_step=True, log_spec=True):
        L = _cabi.lib()
        self.name = name
        self.lam = np.ascontiguousarray(lam, dtype=np.float64)
        self.parnames = tuple(parnames)
        self.kind = kind
        self.log_ids = tuple(log_ids)
        self.log_step = bool(log_step)
        self.log_spec = bool(log_spec)
        self.npix_t = len(self.lam)
        n = self.npix_t
        # knot tables (host, plain C) -> device
        h, hinv = np.zeros(n - 1), np.zeros(n - 1)
        cp, winv = np.zeros(n - 2), np.zeros(n - 2)
        L.rvs_knot_tables(_dev.hptr(self.lam), n, _dev.hptr(h), _dev.hptr(hinv),
                          _dev.hptr(cp), _dev.hptr(winv))
        self.knots = _cabi.Knots()
        st = L.rvs_knot_info(_dev.hptr(self.lam), n, int(self.log_step),
                             ctypes.byref(self.knots))
        if st != 0:
            raise ValueError(f'template wavelength grid of {name} is not uniform (status {st})')
        self._tabs = [_dev.upload(_, np.float64) for _ in (self.lam, h, hinv, cp, winv)]
        (self.knots.d_lam_t, self.knots.d_h, self.knots.d_hinv, self.knots.d_cp,
         self.knots.d_winv) = [t.data_ptr() for t in self._tabs]
        # grid rows, padded to a 16-byte multiple
        dats = np.asarray(dats)
        self.box = None
        dense = False
        if kind == 'regulargrid':
            # A grid without missing nodes is stored in HBM in the C order of its node
            # table (row id = flat grid position), whatever order the product lists its
            # nodes in: the 2^d corner rows of a cell then form a box the copy engine can
            # fetch as one tensor tile (rvs_gridbox).  Node ids are internal to the bank,
            # so the permutation is applied to rows, node coordinates and id table alike.
            idg = np.asarray(idgrid)
            flat = idg.reshape(-1)
            dense = bool(flat.size == dats.shape[0] and (flat >= 0).all()
                         and np.array_equal(np.sort(flat), np.arange(flat.size)))
            if dense and not np.array_equal(flat, np.arange(flat.size)):
                dats = dats[flat]
                vecs = np.asarray(vecs)[:, flat]
                idgrid = np.arange(flat.size).reshape(idg.shape)
        if dats.dtype == np.float64:
            d32 = dats.astype(np.float32)
            if np.array_equal(d32.astype(np.float64), dats):
                dats = d32          # lossless: the product was float32 upstream
        self.grid_f64 = int(dats.dtype == np.float64)
        self.ld = (n + 31) // 32 * 32
        torch = _dev.torch_mod()
        self.grid = _dev.zeros((dats.shape[0], self.ld),
                               np.float64 if self.grid_f64 else np.float32)
        self.grid[:, :n] = torch.from_numpy(np.ascontiguousarray(dats)).to(self.grid.device)
        self.nnode = dats.shape[0]
        if kind == 'regulargrid':
            self.uvecs = [np.asarray(_, dtype=np.float64) for _ in uvecs]
            self.idgrid = np.asarray(idgrid)
            self.ndim = len(self.uvecs)
            self.lens = np.array([len(_) for _ in self.uvecs])
            self.corners = np.array(list(itertools.product([0, 1], repeat=self.ndim)))
            self.nvert = 2**self.ndim
            vecs = np.asarray(vecs, dtype=np.float64)
            self.ptp = np.ptp(vecs, axis=1)
            self.tree = scipy.spatial.cKDTree(vecs.T / self.ptp[None, :])
            self.gridmap = None
            # banks with the same signature locate a parameter vector identically
            # (same nodes, same id table): one rvs_locate_grid call serves them all
            import hashlib
            hsh = hashlib.sha1()
            for arr in list(self.uvecs) + [np.ascontiguousarray(self.idgrid, dtype=np.int64),
                                           np.ascontiguousarray(vecs)]:
                hsh.update(np.ascontiguousarray(arr).tobytes())
            self.locate_signature = (hsh.hexdigest(), self.log_ids)
            if self.ndim <= 5:      # device-side vertex location (rvs_locate_grid)
                uoff = np.concatenate([[0], np.cumsum(self.lens)])
                self._gm = (_dev.upload(np.concatenate(self.uvecs), np.float64),
                            _dev.upload(self.idgrid.reshape(-1), np.int32),
                            _dev.upload(vecs.T / self.ptp[None, :], np.float64))
                gm = _cabi.GridMap()
                gm.d_uvec, gm.d_idgrid = self._gm[0].data_ptr(), self._gm[1].data_ptr()
                gm.d_vnorm, gm.nnode = self._gm[2].data_ptr(), vecs.shape[1]
                gm.ndim = self.ndim
                for i in range(self.ndim):
                    gm.ptp[i] = float(self.ptp[i])
                for i in range(self.ndim):
                    gm.len[i], gm.uoff[i] = int(self.lens[i]), int(uoff[i])
                self.gridmap = gm
            if dense and self.ndim == 4 and not self.grid_f64 and \
                    not os.environ.get('RVS_NO_TMA'):
                box = _cabi.GridBox()
                lens = (ctypes.c_int32 * 4)(*[int(_) for _ in self.lens])
                rc = L.rvs_gridbox_init(ctypes.byref(box), self.grid.data_ptr(), self.ld, 4,
                                        ctypes.cast(lens, ctypes.c_void_p))
                _cabi.check(rc, 'rvs_gridbox_init')
                self.box = box
        elif kind == 'triangulation':
            self.triang = triang
            self.ndim = triang.ndim
            self.nvert = self.ndim + 1
            self.extraflags = np.asarray(extraflags, dtype=np.float64).reshape(-1)
        else:
            raise RuntimeError('Unrecognized interpolation type ' + str(kind))

    def tapcap(self, vsini_max):
        """One-sided length of the rotation kernel at vsini_max (+1), as
        rvs_chisq_fused sizes it (spec_fit.py:666-667,590)."""
        if not vsini_max > 0:
            return 0
        return int(np.ceil(vsini_max / 299792.458 / self.knots.lnstep + 1)) + 1

    # ---------------------------------------------------------- host: vertices
    def locate(self, params):
        """(ids int32 (K,nvert), w float64 (K,nvert), outside float64 (K,)).
        outside: 0 inside; >0 off-grid measure; NaN = no template (Delaunay
        point outside the hull)."""
        q = map_params(params, self.log_ids)
        K = q.shape[0]
        ids = np.zeros((K, self.nvert), dtype=np.int32)
        w = np.zeros((K, self.nvert))
        outside = np.zeros(K)
        if self.kind == 'regulargrid':
            nd = self.ndim
            pos = np.empty((K, nd), dtype=np.int64)
            for i in range(nd):
                pos[:, i] = np.searchsorted(self.uvecs[i], q[:, i], 'right') - 1
            out = np.any((pos < 0) | (pos >= self.lens[None, :] - 1), axis=1)
            inn = np.nonzero(~out)[0]
            if len(inn):
                pc = pos[inn][:, None, :] + self.corners[None, :, :]      # (k,16,nd)
                cid = self.idgrid[tuple(pc[..., i] for i in range(nd))]   # (k,16)
                hole = np.any(cid < 0, axis=1)
                x = np.empty((len(inn), nd))
                for i in range(nd):
                    u = self.uvecs[i]
                    pi = pos[inn, i]
                    x[:, i] = (q[inn, i] - u[pi]) / (u[pi + 1] - u[pi])
                ww = np.prod(np.where(self.corners[None, :, :] == 1, x[:, None, :],
                                      1 - x[:, None, :]), axis=2)
                good = inn[~hole]
                ids[good] = cid[~hole]
                w[good] = ww[~hole]
                out[inn[hole]] = True
            bad = np.nonzero(out)[0]
            if len(bad):
                fin = np.isfinite(q[bad]).all(axis=1)
                w[bad, 0] = 1
                if self.nvert > 1:
                    ids[bad, 1] = -1     # single-row item (see rvs_b200.h)
                if fin.any():
                    dist, near = self.tree.query(q[bad[fin]] / self.ptp[None, :])
                    ids[bad[fin], 0] = near
                    outside[bad[fin]] = dist
                # non-finite mapped parameters (teff <= 0): first node
                # (spec_inter.py:156-159); no finite off-grid measure exists
                outside[bad[~fin]] = np.nan
        else:
            nd = self.ndim
            with np.errstate(invalid='ignore'):
                sx = self.triang.find_simplex(np.where(np.isfinite(q), q, 1e300))
            ok = sx >= 0
            if ok.any():
                T = self.triang.transform[sx[ok]]                       # (k, nd+1, nd)
                b = np.einsum('kij,kj->ki', T[:, :nd, :], q[ok] - T[:, nd, :])
                b = np.concatenate([b, 1 - b.sum(axis=1, keepdims=True)], axis=1)
                vid = self.triang.simplices[sx[ok]]
                ids[ok] = vid
                w[ok] = b
                outside[ok] = (self.extraflags[vid] * b).sum(axis=1)
            outside[~ok] = np.nan
            w[~ok, 0] = 1
        return ids, w, outside

    # ------------------------------------------------------------- device ops
    def build(self, ids, w, vsini=None, out=None, status=None):
        """rvs_template_build for K items -> (yz tensor (K, npix_t, 2), status)."""
        K = ids.shape[0]
        d_ids = _dev.upload(ids, np.int32)
        d_w = _dev.upload(w, np.float64)
        d_vs = None if vsini is None else _dev.upload(vsini, np.float64)
        if out is None:
            out = _dev.empty((K, self.npix_t, 2), np.float64)
        if status is None:
            status = _dev.empty((K,), np.int32)
        rc = _cabi.lib().rvs_template_build(
            _dev.ptr(self.grid), self.grid_f64, self.ld, ctypes.byref(self.knots),
            _dev.ptr(d_ids), _dev.ptr(d_w), self.nvert, _dev.ptr(d_vs), int(self.log_spec),
            K, _dev.ptr(out), self.npix_t, _dev.ptr(status), _dev.stream())
        _cabi.check(rc, 'rvs_template_build')
        return out, status

    def template(self, params, vsini=None):
        """Host copy of interpolated (and optionally broadened) templates for
        parameter vectors (K, ndim): (spec (K, npix_t), outside (K,))."""
        params = np.array(params, dtype=np.float64, ndmin=2)
        ids, w, outside = self.locate(params)
        vs = None if vsini is None else np.broadcast_to(
            np.asarray(vsini, dtype=np.float64), (len(params),))
        yz, _ = self.build(ids, w, vs)
        spec = _dev.download(yz[:, :, 0])
        spec[np.isnan(outside) & (self.kind == 'triangulation')] = np.nan
        return spec, outside


class SpecInterpolator:
    """Same role and method names as reference spec_inter.py:197-286."""

    def __init__(self, bank):
        self.bank = bank
        self.name = bank.name
        self.lam = bank.lam
        self.parnames = bank.parnames
        self.log_step = bank.log_step

    def _vec(self, param0):
        if isinstance(param0, dict):
            try:
                param0 = [param0[_] for _ in self.parnames]
            except KeyError as exc:
                raise ValueError(f'The parameter {exc.args[0]} not found. '
                                 'Required list of parameters is: ' +
                                 ','.join(self.parnames))
        return np.asarray(param0, dtype=np.float64)

    def outsideFlag(self, param0):
        return float(self.bank.locate(self._vec(param0)[None, :])[2][0])

    def eval(self, param0):
        spec, outside = self.bank.template(self._vec(param0)[None, :])
        if self.bank.kind == 'triangulation' and np.isnan(outside[0]):
            return np.nan
        return spec[0]


class interp_cache:
    """Registry of interpolators (reference spec_inter.py:289-293)."""
    interps = {}
    template_lib = None


def clear_caches():
    """Forget every cached interpolator, CCF bank and likelihood engine."""
    import sys
    interp_cache.interps.clear()
    interp_cache.template_lib = None
    for mod, attr in (('fitter_ccf', 'CCFCache'), ('spec_fit', '_engine_cache')):
        m = sys.modules.get(__package__ + '.' + mod)
        if m is not None:
            obj = getattr(m, attr)
            (obj.banks if attr == 'CCFCache' else obj).clear()


def register_bank(bank, template_lib=None):
    """Put an in-memory bank into the registry `getInterpolator` serves."""
    if template_lib is not None:
        interp_cache.template_lib = template_lib
    it = SpecInterpolator(bank)
    interp_cache.interps[bank.name] = it
    return it


def bank_from_setup(setup, kind='regulargrid', name=None):
    """TemplateBank from a synth.make_setup() product (tests / bench)."""
    name = name or setup['name']
    if kind == 'regulargrid':
        return TemplateBank(name, setup['lam'], setup['dats'], setup['parnames'],
                            kind=kind, uvecs=setup['uvecs'], idgrid=setup['idgrid'],
                            vecs=setup['vec'], log_step=setup['log_step'])
    tri, dats64, flags = build_triangulation(setup['vec'], setup['dats'])
    return TemplateBank(name, setup['lam'], dats64, setup['parnames'], kind='triangulation',
                        triang=tri, extraflags=flags, log_step=setup['log_step'])


def build_triangulation(vec_mapped, dats):
    """Delaunay product as the reference's make_nd.py:101-140 lays it out
    (seeded 1e-6 perturbation, 2^d ghost vertices at the bounding box +-20 %
    carrying nearest-neighbour spectra and flag 1).  Offline preparation, kept
    here only so that in-memory banks can be made without HDF5."""
    vec = np.asarray(vec_mapped, dtype=np.float64)
    state = np.random.get_state()
    np.random.seed(1)
    vec = vec + np.random.uniform(-1e-6, 1e-6, size=vec.shape)
    np.random.set_state(state)
    nd = vec.shape[0]
    span = np.ptp(vec, axis=1)
    lo, hi = vec.min(axis=1) - 0.2 * span, vec.max(axis=1) + 0.2 * span
    ghosts = np.array([[(hi[j] if (i >> j) & 1 else lo[j]) for j in range(nd)]
                       for i in range(2**nd)]).T
    near = scipy.spatial.cKDTree(vec.T).query(ghosts.T)[1]
    nspec = dats.shape[0]
    allvec = np.hstack((vec, ghosts))
    alld = np.append(dats, dats[near], axis=0).astype(np.float64)
    flags = np.concatenate((np.zeros(nspec), np.ones(ghosts.shape[1])))
    return scipy.spatial.Delaunay(allvec.T), alld, flags


def getInterpolator(HR, config, warmup_cache=False, cache=None):
    """reference spec_inter.py:296-398.  Banks registered in memory are served
    as is; otherwise the reference's on-disk products are loaded (needs h5py)."""
    if cache is None:
        cache = interp_cache.interps
        lib = config['template_lib']
        if interp_cache.template_lib is not None and lib != interp_cache.template_lib:
            # another template library: nothing cached for the old one may be served
            # (reference spec_inter.py:321-324) -- interpolators, CCF banks and the
            # engines that captured them
            clear_caches()
            interp_cache.template_lib = lib
    if HR not in cache:
        from . import bank_io
        cache[HR] = SpecInterpolator(bank_io.load_bank(HR, config))
        interp_cache.template_lib = config['template_lib']
    return cache[HR]


def getSpecParams(setup, config):
    """reference spec_inter.py:401-417."""
    return getInterpolator(setup, config).parnames
