"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

CPU restatement (numpy + the C functions of oracle/rvs_oracle.c) of the
rvspecfit per-spectrum likelihood hot path: template interpolation, vsini
broadening, Doppler spline resampling, continuum-marginalised chi-square, the
RV-grid scan statistics, the fit driver and the CCF first guess.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.  rvspecfit_b200/ never does.

Parity status: PINNED for everything except the Hessian-based outputs.
tests/test_oracle.py checks this module against tests/golden/*.npz, which were
produced by running the reference package itself (tests/golden/make_golden.py)
on seeded inputs.  UNPINNED: `param_err`, `param_covar`, `bad_hessian` of
`process` -- the reference computes them with numdifftools (unpinned dependency,
pyproject.toml:24; call sites vel_fit.py:713-716), which is absent here; this
module uses `central_hessian` below instead and both sides of the GPU parity
test use that same routine.

Every function cites the reference file:line it follows
(paths relative to /root/reference/py/rvspecfit/).
"""
import ctypes
import itertools
import math
import os

import numpy as np
import scipy.linalg
import scipy.optimize
import scipy.signal
import scipy.spatial
import scipy.interpolate
import scipy.stats

C_KMS = 299792.458  # spec_fit.py:23

_here = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)


def _load(path):
    if not os.path.exists(path):
        raise RuntimeError(f'{path} missing: run `make -C oracle`')
    return ctypes.CDLL(path)


_lib = None


def clib():
    global _lib
    if _lib is None:
        _lib = _load(os.path.join(_here, 'librvs_oracle.so'))
        _lib.orc_spline_construct.argtypes = [_dp, _dp, ctypes.c_int] + [_dp] * 5
        _lib.orc_spline_construct.restype = None
        _lib.orc_spline_eval.argtypes = [_dp, ctypes.c_int, ctypes.c_int] + \
            [_dp] * 5 + [ctypes.c_int, _dp]
        _lib.orc_spline_eval.restype = ctypes.c_int
        _lib.orc_chisq0_chol.argtypes = [_dp] * 4 + [ctypes.c_int, ctypes.c_int, _dp, _dp]
        _lib.orc_chisq0_chol.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


# ---------------------------------------------------------------- spline
class Spline:
    """Natural cubic spline on a uniform (lin or log) knot grid.
    Follows spliner.py:8-53 over spliner.c:7-108."""

    def __init__(self, xs, ys, log_step=True, lib=None, names=None):
        self.lib = lib or clib()
        self.names = names or ('orc_spline_construct', 'orc_spline_eval')
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        ys = np.ascontiguousarray(ys, dtype=np.float64)
        n = len(xs)
        self.xs, self.n, self.log_step = xs, n, int(log_step)
        self.A, self.B, self.C, self.D, self.h = [np.zeros(n - 1) for _ in range(5)]
        getattr(self.lib, self.names[0])(_p(xs), _p(ys), n, _p(self.A), _p(self.B),
                                         _p(self.C), _p(self.D), _p(self.h))

    def __call__(self, ex):
        ex = np.ascontiguousarray(ex, dtype=np.float64)
        out = np.zeros(len(ex))
        fn = getattr(self.lib, self.names[1])
        if self.names[1] == 'evaler':   # reference signature carries hs too
            st = fn(_p(ex), len(ex), self.n, _p(self.xs), _p(self.h), _p(self.A),
                    _p(self.B), _p(self.C), _p(self.D), self.log_step, _p(out))
        else:
            st = fn(_p(ex), len(ex), self.n, _p(self.xs), _p(self.A), _p(self.B),
                    _p(self.C), _p(self.D), self.log_step, _p(out))
        if st != 0:
            raise AssertionError(f'spline evaluation status {st}')  # spliner.py:51
        return out


def reference_spline_lib():
    """The reference's own spliner.c compiled by oracle/Makefile (or None)."""
    path = os.path.join(_here, '_ref', 'libspliner_ref.so')
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    lib.construct.argtypes = [_dp, _dp, ctypes.c_int] + [_dp] * 5
    lib.construct.restype = None
    lib.evaler.argtypes = [_dp, ctypes.c_int, ctypes.c_int] + [_dp] * 6 + [ctypes.c_int, _dp]
    lib.evaler.restype = ctypes.c_int
    return lib


# ---------------------------------------------------------- interpolators
def map_params(p, log_ids=(0,)):
    """read_grid.py:127-145 (LogParamMapper.forward)."""
    q = np.array(p, dtype=np.float64)
    for i in log_ids:
        q[i] = np.log10(q[i])
    return q


class PolylinearGrid:
    """Regular-grid corner-weighted interpolation and its off-grid measure.
    Follows spec_inter.py:95-194 (GridInterp) and :62-92 (GridOutsideCheck)."""

    def __init__(self, uvecs, idgrid, vecs, dats, exp=True):
        self.uvecs = [np.asarray(_, dtype=np.float64) for _ in uvecs]
        self.idgrid, self.dats, self.exp = idgrid, dats, exp
        self.ndim = len(uvecs)
        self.lens = np.array([len(_) for _ in uvecs])
        self.corners = np.array(list(itertools.product([0, 1], repeat=self.ndim)))
        self.ptp = np.ptp(vecs, axis=1)
        self.tree = scipy.spatial.cKDTree(vecs.T / self.ptp[None, :])

    def _cell(self, q):
        # np.digitize(x, bins) - 1 == searchsorted(bins, x, 'right') - 1
        return np.array([np.searchsorted(self.uvecs[i], q[i], 'right') - 1
                         for i in range(self.ndim)])

    def vertices(self, q):
        """(node ids, weights, inside?) for mapped parameter vector q."""
        q = np.asarray(q, dtype=np.float64)
        cell = self._cell(q)
        if np.any((cell < 0) | (cell >= self.lens - 1)):
            if not np.isfinite(q).all():          # spec_inter.py:156-159
                return np.array([0]), np.array([1.]), False
            return np.array([self.tree.query(q / self.ptp)[1]]), np.array([1.]), False
        ids = self.idgrid[tuple((cell[None, :] + self.corners).T)]
        if np.any(ids < 0):                        # hole in the grid
            return np.array([self.tree.query(q / self.ptp)[1]]), np.array([1.]), False
        x = np.array([(q[i] - self.uvecs[i][cell[i]]) /
                      (self.uvecs[i][cell[i] + 1] - self.uvecs[i][cell[i]])
                      for i in range(self.ndim)])
        w = np.prod(np.where(self.corners == 1, x[None, :], 1 - x[None, :]), axis=1)
        return ids, w, True

    def __call__(self, q):
        ids, w, inside = self.vertices(q)
        if inside:
            val = np.dot(w, self.dats[ids, :])     # f32 rows promoted to f64
        else:
            val = self.dats[ids[0]]
        return np.exp(val) if self.exp else val

    def outside(self, q):
        q = np.asarray(q, dtype=np.float64)
        cell = self._cell(q)
        out = bool(np.any((cell < 0) | (cell >= self.lens - 1)))
        if not out:
            out = bool((self.idgrid[tuple((cell[None, :] + self.corners).T)] == -1).any())
        if out:
            return self.tree.query(q / self.ptp)[0]
        return 0


def ghost_vertices(vec, pad=0.2):
    """make_nd.py:18-52: corners of the bounding box grown by `pad`."""
    ndim = vec.shape[0]
    span = np.ptp(vec, axis=1)
    lo, hi = vec.min(axis=1) - pad * span, vec.max(axis=1) + pad * span
    pts = [[(hi[j] if (i >> j) & 1 else lo[j]) for j in range(ndim)]
           for i in range(2**ndim)]
    return np.array(pts).T


def build_triangulation(vec_mapped, dats):
    """The Delaunay product of make_nd.py:101-140: seeded 1e-6 perturbation,
    2^d ghost vertices carrying nearest-neighbour spectra and flag 1."""
    vec = vec_mapped.astype(float)
    st = np.random.get_state()
    np.random.seed(1)
    vec = vec + np.random.uniform(-1e-6, 1e-6, size=vec.shape)
    np.random.set_state(st)
    ghosts = ghost_vertices(vec)
    near = scipy.spatial.cKDTree(vec.T).query(ghosts.T)[1]
    nspec = dats.shape[0]
    vec = np.hstack((vec, ghosts))
    dats64 = np.append(dats, dats[near], axis=0).astype(np.float64)
    flags = np.concatenate((np.zeros(nspec), np.ones(ghosts.shape[1])))[:, None]
    tri = scipy.spatial.Delaunay(vec.T)
    return tri, dats64, flags, vec


class SimplexInterp:
    """Barycentric interpolation over a Delaunay triangulation
    (spec_inter.py:11-59, TriInterp)."""

    def __init__(self, tri, dats, exp=True):
        self.tri, self.dats, self.exp = tri, dats, exp
        self.ndim = tri.ndim

    def vertices(self, q):
        q = np.asarray(q, dtype=np.float64)
        s = int(self.tri.find_simplex(q))
        if s == -1:
            return None, None
        d = self.ndim
        T = self.tri.transform[s]
        b = np.empty(d + 1)
        b[:d] = T[:d, :].dot(q - T[d, :])
        b[d] = 1 - b[:d].sum()
        return self.tri.simplices[s], b

    def __call__(self, q):
        ids, b = self.vertices(q)
        if ids is None:
            return np.nan
        val = (self.dats[ids, :] * b[:, None]).sum(axis=0)
        if self.exp:
            val = np.exp(val)
        if val.size == 1:
            val = float(val[0])
        return val


class Interpolator:
    """spec_inter.py:197-286 (SpecInterpolator): mapper + interpolant + off-grid
    measure + wavelength grid."""

    def __init__(self, name, interper, extraper, lam, parnames, log_ids=(0,),
                 log_step=True):
        self.name, self.interper, self.extraper = name, interper, extraper
        self.lam = np.ascontiguousarray(lam, dtype=np.float64)
        self.parnames, self.log_ids, self.log_step = tuple(parnames), log_ids, log_step

    def outsideFlag(self, p):
        return self.extraper(map_params(p, self.log_ids))

    def eval(self, p):
        if isinstance(p, dict):
            p = [p[_] for _ in self.parnames]
        return self.interper(map_params(p, self.log_ids))


REGISTRY = {}


def register_setup(setup, kind='grid'):
    """Register a synth.make_setup() product under setup['name']."""
    if kind == 'grid':
        g = PolylinearGrid(setup['uvecs'], setup['idgrid'], setup['vec'], setup['dats'])
        it = Interpolator(setup['name'], g, g.outside, setup['lam'], setup['parnames'])
    else:
        tri, d64, flags, _ = build_triangulation(setup['vec'], setup['dats'])
        it = Interpolator(setup['name'], SimplexInterp(tri, d64, True),
                          SimplexInterp(tri, flags, False), setup['lam'],
                          setup['parnames'])
    REGISTRY[setup['name']] = it
    return it


# ---------------------------------------------------------------- vsini
def _rot_primitives(x, eps):
    """spec_fit.py:495-547: primitives of K(x) and x K(x),
    K ~ c1 sqrt(1-x^2) + c2 (1-x^2)."""
    x = np.clip(x, -1.0, 1.0)
    nrm = np.pi * (1 - eps / 3.0)
    c1 = 2 * (1 - eps) / nrm
    c2 = (np.pi / 2.0) * eps / nrm
    u = 1 - x**2
    root = np.sqrt(u)
    k0 = c1 * (0.5 * (x * root + np.arcsin(x))) + c2 * (x - (x**3) / 3.0)
    k1 = c1 * (-1.0 / 3.0 * u * root) + c2 * ((x**2) / 2.0 - (x**4) / 4.0)
    return k0, k1


def _rot_segment(xa, xb, slope, icpt, eps):
    """spec_fit.py:550-562."""
    k0b, k1b = _rot_primitives(xb, eps)
    k0a, k1a = _rot_primitives(xa, eps)
    return slope * (k1b - k1a) + icpt * (k0b - k0a)


def vsini_kernel(R, eps=0.6):
    """spec_fit.py:565-625: overlap weights of the rotation profile with the
    triangular pixel basis; length 2*ceil(R+1)+1, symmetric, unit sum."""
    assert R > 0
    kmax = int(np.ceil(R + 1))
    k = np.arange(0, kmax + 1)
    w = np.zeros(len(k))
    lo, hi = np.clip(k / R, -1, 1), np.clip((k + 1) / R, -1, 1)
    m = hi > lo
    if m.any():
        w[m] += _rot_segment(lo[m], hi[m], -R, 1 + k[m], eps)
    lo, hi = np.clip((k - 1) / R, -1, 1), np.clip(k / R, -1, 1)
    m = hi > lo
    if m.any():
        w[m] += _rot_segment(lo[m], hi[m], R, 1 - k[m], eps)
    full = np.concatenate([w[:0:-1], w])
    return full / full.sum()


def rotational_broaden(lam_t, templ, vsini, eps=0.6):
    """spec_fit.py:628-682 (convolve_vsini)."""
    if vsini <= 0:
        return templ.copy()
    ratios = lam_t[1:] / lam_t[:-1]
    assert np.allclose(ratios, ratios[0])
    R = (vsini / C_KMS) / np.log(ratios[0])
    if R < 1e-9:
        return templ.copy()
    return scipy.signal.convolve(templ, vsini_kernel(R, eps), mode='same',
                                 method='auto')


# ------------------------------------------------------------ likelihood
class SpecData:
    """spec_fit.py:70-145."""

    def __init__(self, name, lam, spec, espec, badmask=None, resolution=None):
        self.name = name
        self.resolution = resolution      # ResolMatrix or None (spec_fit.py:54-67,111)
        self.lam = np.ascontiguousarray(lam, dtype=np.float64)
        self.spec = np.ascontiguousarray(spec, dtype=np.float64)
        self.espec = np.ascontiguousarray(espec, dtype=np.float64)
        self.badmask = (np.zeros(len(self.spec), dtype=bool)
                        if badmask is None else np.asarray(badmask))
        self._basis = {}


class ResolMatrix:
    """spec_fit.py:54-67: holder of a (sparse) resolution matrix."""

    def __init__(self, mat):
        self.mat = mat


def construct_resol_mat(lam, resol=None, width=None):
    """spec_fit.py:410-468: banded matrix of Gaussian line-spread functions
    (sigma = lam/R/2.35 or `width`), truncated at 5 sigma, every COLUMN
    normalised to unit sum, stored by diagonals."""
    import scipy.sparse
    assert (resol is None) != (width is None)
    lam = np.asarray(lam, dtype=np.float64)
    n = len(lam)
    if resol is not None:
        sig = lam / resol / 2.35
    else:
        sig = np.zeros(n) + width
    assert np.all(np.diff(lam) > 0)
    i1 = np.maximum(np.searchsorted(lam, lam - 5 * sig, 'left'), 0)
    i2 = np.minimum(np.searchsorted(lam, lam + 5 * sig, 'right'), n - 1)
    pix = np.arange(n)
    maxl = min(n, max(np.max(i2 - pix), np.max(pix - i1)))
    offs = np.arange(-maxl, maxl + 1)
    nb = pix[None, :] + offs[:, None]              # neighbour of column j on diagonal row k
    ok = (nb >= 0) & (nb < n)
    nb[~ok] = 0
    X = np.exp(-0.5 * ((lam[nb] - lam[None, :]) / sig[None, :])**2) * ok
    X = X / X.sum(axis=0)[None, :]
    # X[k, j] = weight of pixel j+offs[k] in the profile centred on j; spdiags wants
    # data[k, c] = M[c-offs[k], c]: the reference's index shuffle (spec_fit.py:463-465)
    yid = (pix[None, :] + (n - offs)[:, None]) % n
    xid = yid * 0 + maxl + offs[:, None]
    return ResolMatrix(scipy.sparse.spdiags(X[xid, yid], offs, n, n))


def convolve_resol(spec, resol_matrix):
    """spec_fit.py:471-489."""
    return resol_matrix.mat @ spec


def continuum_basis(lam, npoly, rbf=True):
    """spec_fit.py:148-176."""
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


def _basis(sd, npoly, rbf):
    key = (npoly, rbf)
    if key not in sd._basis:
        sd._basis[key] = continuum_basis(sd.lam, npoly, rbf)
    return sd._basis[key]


def marginal_chisq_svd(spec, templ, polys, espec, get_coeffs=False):
    """spec_fit.py:255-303."""
    D = spec / espec
    G = (templ / espec)[None, :] * polys
    v = G @ D[:, None]
    M = np.dot(G, G.T)
    u, s, vt = scipy.linalg.svd(M, check_finite=False)
    a = vt.T @ ((1. / s)[:, None] * u.T) @ v
    chisq = np.sum(np.log(s)) + 2 * np.log(espec).sum() + \
        np.linalg.norm(D - a.T @ G)**2
    return (chisq, a.flatten()) if get_coeffs else chisq


def marginal_chisq(spec, templ, polys, espec, get_coeffs=False):
    """spec_fit.py:306-354: Cholesky route unless coefficients are wanted or it
    fails / is not finite; SVD route otherwise."""
    if not get_coeffs:
        out = ctypes.c_double()
        spec, templ, espec = [np.ascontiguousarray(_, dtype=np.float64)
                              for _ in (spec, templ, espec)]
        polys = np.ascontiguousarray(polys)
        st = clib().orc_chisq0_chol(_p(spec), _p(templ), _p(polys), _p(espec),
                                    len(spec), polys.shape[0], ctypes.byref(out), None)
        if st == 0:
            return out.value
    return marginal_chisq_svd(spec, templ, polys, espec, get_coeffs)


class TemplateCache:
    """Small LRU like functools.lru_cache(100) on getCurTempl (spec_fit.py:357)
    plus the spline cache of find_best (spec_fit.py:1060)."""

    def __init__(self, n=100):
        self.n, self.d = n, {}

    def get(self, key, make):
        if key in self.d:
            val = self.d.pop(key)
        else:
            val = make()
            if len(self.d) >= self.n:
                self.d.pop(next(iter(self.d)))
        self.d[key] = val
        return val


_templates = TemplateCache(100)


def current_template(setup, atm, rot):
    """spec_fit.py:357-407 (getCurTempl) without the random tag: returns
    (outside, lam, spec, interpolator)."""
    def make():
        it = REGISTRY[setup]
        outside = float(it.outsideFlag(atm))
        spec = np.ascontiguousarray(it.eval(atm), dtype=np.float64)
        if spec.ndim == 0:
            spec = np.full(len(it.lam), float(spec))
        if outside > 0:
            mx = np.abs(spec).max()
            if mx > 1e100 or not np.isfinite(mx):
                outside = np.nan
        if np.isfinite(outside) and rot is not None:
            spec = rotational_broaden(it.lam, spec, *rot)
        return [outside, it.lam, spec, it, None]
    return _templates.get((setup, tuple(atm), None if rot is None else tuple(rot)), make)


def check_overlap(t0, t1, s0, s1, vmin, vmax):
    """spec_fit.py:786-794."""
    for v in (vmin, vmax):
        k = np.sqrt((1 + v / C_KMS) / (1 - v / C_KMS))
        if t0 * k > s0 or t1 * k < s1:
            raise RuntimeError(f'template ({t0},{t1}) does not cover ({s0},{s1}) '
                               f'for velocities {vmin} {vmax}')


def resample(spl, vel, lam):
    """spec_fit.py:707-727 (evalRV)."""
    beta = vel / C_KMS
    return spl(lam * np.sqrt((1 - beta) / (1 + beta)))


def get_chisq(specdata, vel, atm, rot=None, options=None, config=None,
              full_output=False, espec_systematic=None, outside_penalty=True,
              fast_interp=False, resol_params=None):
    """spec_fit.py:797-989.  fast_interp: nearest-knot lookup instead of the
    spline (spec_fit.py:913-918).  resol_params (dictionary by setup) or
    SpecData.resolution: the resampled template is multiplied by the resolution
    matrix before the continuum fit (spec_fit.py:922-929)."""
    npoly = options.get('npoly') or 5
    rbf = options.get('rbf_continuum', True)
    acc = 0
    bad = 10 * sum(len(_.lam) for _ in specdata)
    rot = None if rot is None else tuple(rot)
    atm = tuple(atm)
    models, raws, chis, reds, npixs = [], [], [], [], []
    for sd in specdata:
        ent = current_template(sd.name, atm, rot)
        outside, tlam, tspec, it = ent[:4]
        if not np.isfinite(outside):
            acc += 1000 * bad
            chis.append(np.nan)
            reds.append(np.nan)
            models.append(np.zeros(len(sd.lam)) + np.nan)
            continue
        if outside_penalty:
            acc += outside * bad
        check_overlap(tlam[0], tlam[-1], sd.lam[0], sd.lam[-1],
                      min(config['min_vel'], vel), max(config['max_vel'], vel))
        if fast_interp:
            beta = vel / C_KMS
            ev = tspec[np.searchsorted(tlam, np.sqrt((1 - beta) / (1 + beta)) * sd.lam)]
        else:
            if ent[4] is None:
                ent[4] = Spline(tlam, tspec, log_step=it.log_step)
            ev = resample(ent[4], vel, sd.lam)
        if resol_params is not None:
            ev = convolve_resol(ev, resol_params[sd.name])
        if getattr(sd, 'resolution', None) is not None:
            if resol_params is not None:
                raise ValueError('You are not allowed to set resol_param together with'
                                 'the resolution of each SpecData')
            ev = convolve_resol(ev, sd.resolution)
        polys = _basis(sd, npoly, rbf)
        if espec_systematic is not None:
            sy = espec_systematic[sd.name] if isinstance(espec_systematic, dict) \
                else espec_systematic
            es = np.sqrt(sy**2 + sd.espec**2)
        else:
            es = sd.espec
        cur = marginal_chisq(sd.spec, ev, polys, es, get_coeffs=full_output)
        if full_output:
            cur, co = cur
            mod = np.dot(co, polys * ev)
            raws.append(ev)
            models.append(mod)
            dev = (mod - sd.spec) / sd.espec
            good = ~sd.badmask
            chis.append(np.sum(dev[good]**2))
            npixs.append(good.sum())
            reds.append(chis[-1] / npixs[-1])
        if not np.isfinite(float(cur)):
            if outside > 0 and np.isfinite(ev).all():
                continue
            raise RuntimeError('The log(likelihood) value is not finite')
        acc += float(cur)
    if full_output:
        return dict(chisq=acc, logl=-0.5 * acc, chisq_array=chis,
                    red_chisq_array=reds, npix_array=npixs, models=models,
                    raw_models=raws)
    return acc


def get_chisq_continuum(specdata, options=None):
    """spec_fit.py:739-783."""
    npoly = options.get('npoly') or 5
    rbf = options.get('rbf_continuum', True)
    ca, ra = np.zeros(len(specdata)), np.zeros(len(specdata))
    for i, sd in enumerate(specdata):
        polys = _basis(sd, npoly, rbf)
        templ = np.ones(len(sd.spec))
        _, co = marginal_chisq(sd.spec, templ, polys, sd.espec, get_coeffs=True)
        dev = (np.dot(co, polys * templ) - sd.spec) / sd.espec
        good = ~sd.badmask
        ca[i] = np.sum(dev[good]**2)
        ra[i] = ca[i] / good.sum()
    return dict(chisq_array=ca, redchisq_array=ra)


def parabola_vertex(x, y, i):
    """spec_fit.py:992-1015."""
    if i == 0 or i == len(x) - 1:
        return x[i]
    a2, a1, _ = np.polyfit(x[i - 1:i + 2], y[i - 1:i + 2], 2)
    val = -a1 / 2 / a2
    assert x[i - 1] < val < x[i + 1]
    return val


def scan_statistics(vel_grid, chisq, quadratic=True):
    """spec_fit.py:1072-1092: argmin, posterior moments on the grid."""
    i1, i2 = np.unravel_index(np.argmin(chisq), chisq.shape)
    pr = np.exp(-0.5 * (chisq[:, i2] - chisq[i1, i2]))
    pr = pr / pr.sum()
    bv = parabola_vertex(vel_grid, chisq[:, i2], i1) if quadratic else vel_grid[i1]
    err = np.sqrt((pr * (vel_grid - bv)**2).sum())
    if err < 1e-10:
        ku, sk = 0, 0
    else:
        ku = (pr * (vel_grid - bv)**4).sum() / err**4
        sk = (pr * (vel_grid - bv)**3).sum() / err**3
    return dict(best_chi=chisq[i1, i2], best_vel=bv, vel_err=err, ibest=int(i2),
                kurtosis=ku, skewness=sk, probs=pr)


def find_best(specdata, vel_grid, params_list, rot=None, options=None, config=None,
              quadratic=True, return_chisq=False, resol_params=None):
    """spec_fit.py:1018-1092."""
    chisq = np.zeros((len(vel_grid), len(params_list)))
    for j, par in enumerate(params_list):
        for i, v in enumerate(vel_grid):
            chisq[i, j] = get_chisq(specdata, v, par, rot, options=options, config=config,
                                    resol_params=resol_params)
    st = scan_statistics(np.asarray(vel_grid), chisq, quadratic)
    st['best_param'] = params_list[st.pop('ibest')]
    if return_chisq:
        st['chisq'] = chisq
    return st


# ------------------------------------------------------------- fit driver
def firstguess(specdata, options=None, config=None, vsinigrid=(None, 10, 100),
               paramsgrid=None):
    """vel_fit.py:13-94."""
    options = options or {}
    if paramsgrid is None:
        paramsgrid = {'logg': [1, 2, 3, 4, 5], 'teff': [3000, 5000, 8000, 10000],
                      'feh': [-2, -1, 0], 'alpha': [0]}
    names = REGISTRY[specdata[0].name].parnames
    plist = []
    for x in itertools.product(*paramsgrid.values()):
        d = dict(zip(paramsgrid.keys(), x))
        plist.append([d[_] for _ in names])
    vg = np.arange(config['min_vel'], config['max_vel'], config['vel_step0'])
    best = np.inf
    for vs in vsinigrid:
        res = find_best(specdata, vg, plist, rot=None if vs is None else (vs,),
                        options=options, config=config)
        if res['best_chi'] < best:
            out = {k: res['best_param'][i] for i, k in enumerate(names)}
            if vs is not None:
                out['vsini'] = vs
            best = res['best_chi']
    return out


class FitVector:
    """vel_fit.py:97-207 (VSiniMapper + ParamMapper): fitted vector layout is
    [vel, (vsini), free atmospheric parameters...]."""

    def __init__(self, names, start, fixed, max_vsini, fit_vsini):
        self.names, self.start, self.fixed = names, start, fixed
        self.max_vsini, self.fit_vsini = max_vsini, fit_vsini

    def unpack(self, p):
        q = list(p)[::-1]
        r = {'vel': q.pop()}
        pen = 0
        if self.fit_vsini:
            raw = q.pop()
            vs = np.clip(raw, 0, self.max_vsini)
            pen += int(raw < 0) * (vs - raw)**2 + int(raw > self.max_vsini) * (vs - raw)**2
            r['vsini'] = vs
        else:
            r['vsini'] = self.start['vsini'] if 'vsini' in self.fixed else None
        r['rot_params'] = None if r['vsini'] is None else (r['vsini'],)
        r['params'] = [self.start[k] if k in self.fixed else q.pop() for k in self.names]
        assert not q
        r['penalty'] = pen
        return r

    def fitted_names(self):
        return ['vel'] + (['vsini'] if self.fit_vsini else []) + \
            [k for k in self.names if k not in self.fixed]


def objective0(pd, args, outside_penalty=True):
    """vel_fit.py:210-230."""
    c = 0
    pri = args.get('priors')
    if pri is not None:
        for i, k in enumerate(args['mapper'].names):
            if k in pri:
                c += ((pri[k][0] - pd['params'][i]) / pri[k][1])**2
    return c + get_chisq(args['specdata'], pd['vel'], pd['params'], pd['rot_params'],
                         options=args['options'], config=args['config'],
                         outside_penalty=outside_penalty)


def objective(p, args):
    """vel_fit.py:233-257."""
    pd = args['mapper'].unpack(p)
    if pd['vel'] > args['max_vel'] or pd['vel'] < args['min_vel'] or \
            (~np.isfinite(pd['params'])).any():
        return 1e30
    args['nfev'] = args.get('nfev', 0) + 1
    return objective0(pd, args) + pd['penalty']


def simplex_start(best_vel, fixed, names, start, max_vsini, fit_vsini):
    """vel_fit.py:272-312."""
    x0, sd = [best_vel], [5]
    if fit_vsini:
        x0.append(np.clip(start['vsini'], 0, max_vsini))
        sd.append(3)
    for k in names:
        if k not in fixed:
            x0.append(start[k])
            sd.append({'logg': 0.5, 'teff': 300, 'feh': 0.5, 'alpha': 0.25}.get(k) or 0.5)
    x0, sd = np.array(x0), np.array(sd)
    n = len(x0)
    rs = np.random.RandomState(43434)
    simp = np.zeros((n + 1, n))
    simp[0] = x0
    simp[1:] = x0[None, :] + sd[None, :] * rs.normal(size=(n, n))
    return x0, simp


def refine_velocity(func, best_vel, min_vel, max_vel, step0, min_step,
                    crit_ratio=5, goal_width=10):
    """vel_fit.py:358-439 (_minimum_sampler)."""
    step = step0
    for it in range(10):
        grid = np.arange(math.ceil((min_vel - best_vel) / step) * step,
                         max_vel - best_vel, step) + best_vel
        best_vel, err, res = func(grid)
        if step < err / crit_ratio or step < min_step:
            break
        if step > err:
            new_step, width = step / crit_ratio, step * goal_width
        else:
            new_step, width = err / crit_ratio * 0.8, err * goal_width
        min_vel = max(best_vel - width, min_vel)
        max_vel = min(best_vel + width, max_vel)
        step = new_step
    return best_vel, err, res


def initial_inverse_hessian(names):
    """vel_fit.py:442-460."""
    d = np.zeros(len(names)) + 0.1**2
    d[names.index('teff')] = 50**2
    if 'vsini' in names:
        d[names.index('vsini')] = 5**2
    d[0] = 1
    return np.diag(d)


def errors_from_hessian(H):
    """vel_fit.py:463-502."""
    dg = np.diag(H)
    inv_dg = 1. / (dg + (dg == 0))
    inv_dg[dg == 0] = np.inf
    bad = False
    try:
        Hi = scipy.linalg.inv(H)
    except (np.linalg.LinAlgError, ValueError):
        bad = True
        Hi = np.diag(inv_dg)
    e0 = np.array(np.diag(Hi))
    b0, b1 = e0 < 0, inv_dg < 0
    if b0.any():
        bad = True
    s1, s2 = b0 & ~b1, b0 & b1
    e0[s1] = inv_dg[s1]
    e0[s2] = 0
    err = np.sqrt(e0)
    err[s2] = np.nan
    if (~np.isfinite(err)).any():
        bad = True
    return err, Hi, bad


def central_hessian(f, x, steps):
    """Second-order central-difference Hessian with one Richardson step
    (h, h/2).  Stand-in for numdifftools.Hessian (absent here): UNPINNED."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)

    def one(hs):
        H = np.zeros((n, n))
        f0 = f(x)
        for i in range(n):
            ei = np.zeros(n)
            ei[i] = hs[i]
            H[i, i] = (f(x + ei) - 2 * f0 + f(x - ei)) / hs[i]**2
            for j in range(i):
                ej = np.zeros(n)
                ej[j] = hs[j]
                H[i, j] = H[j, i] = (f(x + ei + ej) - f(x + ei - ej) - f(x - ei + ej)
                                     + f(x - ei - ej)) / (4 * hs[i] * hs[j])
        return H
    hs = np.asarray(steps, dtype=np.float64)
    return (4 * one(hs / 2) - one(hs)) / 3


HESS_STEP = {'vsini': 1 / 100, 'logg': 0.1 / 100, 'feh': 0.1 / 100,
             'alpha': .01 / 100, 'teff': 1 / 100, 'vrad': 1 / 100}  # vel_fit.py:705-712


def process(specdata, start, fixParam=None, options=None, config=None, priors=None,
            hessian=True):
    """vel_fit.py:505-737."""
    if isinstance(specdata, SpecData):
        specdata = [specdata]
    min_vel, max_vel = config['min_vel'], config['max_vel']
    step0, max_vsini = config['vel_step0'], config['max_vsini']
    min_step = config['min_vel_step']
    second = config.get('second_minimizer') or False
    options = options or {}
    names = REGISTRY[specdata[0].name].parnames
    fixed = fixParam or []
    curparam = tuple(start[k] for k in names)
    if 'vsini' not in start:
        rot, fit_vsini = None, False
    else:
        rot = (start['vsini'],)
        fit_vsini = 'vsini' not in fixed
    vg = np.arange(min_vel, max_vel, step0)
    res = find_best(specdata, vg, [curparam], rot=rot, options=options, config=config)
    x0, simp = simplex_start(res['best_vel'], fixed, names, start, max_vsini, fit_vsini)
    mapper = FitVector(names, start, fixed, max_vsini, fit_vsini)
    args = dict(min_vel=min_vel, max_vel=max_vel, mapper=mapper, specdata=specdata,
                options=options, config=config, priors=priors)
    ok, it = True, 1
    while True:
        r0 = scipy.optimize.minimize(objective, x0, args=args, method='Nelder-Mead',
                                     options={'fatol': 1e-3, 'xatol': 1e-2,
                                              'initial_simplex': simp,
                                              'maxiter': 10000, 'maxfev': np.inf})
        x0, simp = r0['x'], r0['final_simplex'][0]
        if r0['success']:
            break
        if it == 2:
            ok = False
            break
        it += 1
    if second:
        r = scipy.optimize.minimize(objective, r0['x'], method='BFGS', args=args,
                                    options=dict(hess_inv0=initial_inverse_hessian(
                                        mapper.fitted_names())))
    else:
        r = r0
    bp = mapper.unpack(r['x'])
    ret = {'param': dict(zip(names, bp['params']))}
    if fit_vsini:
        ret['vsini'] = bp['vsini']
    bv = bp['vel']
    if bv > max_vel or bv < min_vel:
        bv = max_vel if bv > max_vel else min_vel

    def scan(grid):
        r1 = find_best(specdata, grid, [bp['params']], rot=bp['rot_params'],
                       options=options, config=config)
        return r1['best_vel'], r1['vel_err'], r1
    bv, verr, r1 = refine_velocity(scan, bv, min_vel, max_vel, step0, min_step)
    ret.update(vel=bv, vel_err=verr, vel_skewness=r1['skewness'],
               vel_kurtosis=r1['kurtosis'])
    outp = get_chisq(specdata, bv, bp['params'], bp['rot_params'], options=options,
                     config=config, full_output=True)
    ret['x_opt'] = np.array(r['x'])
    ret['nfev'] = args.get('nfev', 0)
    if hessian:
        pd = dict(bp)
        pd['params'] = list(bp['params'])

        def hf(p):
            pd['params'][:] = p[:]
            return 0.5 * objective0(pd, args)
        H = central_hessian(hf, [ret['param'][k] for k in names],
                            [HESS_STEP[k] for k in names])
        err, cov, badh = errors_from_hessian(H)
        ret.update(param_err=dict(zip(names, err)), param_covar=cov, bad_hessian=badh)
    ret.update(minimize_success=ok, yfit=outp['models'], raw_models=outp['raw_models'],
               chisq=outp['chisq'], logl=outp['logl'], chisq_array=outp['chisq_array'],
               npix_array=outp['npix_array'])
    return ret


# ------------------------------------------------------------------- CCF
def ccf_config(logl0, logl1, npoints, splinestep=1000, maxcontpts=20):
    """make_ccf.py:67-102."""
    c = dict(logl0=logl0, logl1=logl1, npoints=npoints, continuum=True,
             maxcontpts=maxcontpts)
    if splinestep is None:
        c['continuum'] = False
    else:
        c['splinestep'] = max(splinestep, 3e5 * (np.exp((logl1 - logl0) / maxcontpts) - 1))
    return c


def _cont_resid(p, nodes, lam, spec, espec, model=False):
    """make_ccf.py:154-164."""
    mod = np.exp(np.clip(scipy.interpolate.UnivariateSpline(nodes, p, s=0, k=2)(lam),
                         -100, 100))
    return mod if model else (mod - spec) / espec


def fit_continuum(lam, spec, espec, conf):
    """make_ccf.py:105-151."""
    lo = lam.min()
    dl = np.log(1 + conf['splinestep'] / 3e5)
    N = int(np.ceil(np.log(lam.max() / lo) / dl))
    nodes = lo * np.exp(np.arange(N) * dl)
    edges = lo * np.exp((-0.5 + np.arange(N + 1)) * dl)
    med = np.median(spec)
    if med <= 0:
        med = np.abs(med)
        if med == 0:
            med = 1
    bs = scipy.stats.binned_statistic(lam, spec, 'median', bins=edges)
    p0 = np.log(np.maximum(bs.statistic, 1e-3 * med))
    p0[~np.isfinite(p0)] = np.log(med)
    sol = scipy.optimize.least_squares(
        lambda p: _cont_resid(p, nodes, lam, spec, espec), p0, loss='soft_l1')
    return _cont_resid(sol['x'], nodes, lam, spec, espec, model=True)


def fill_masked(lam, spec, bad):
    """make_ccf.py:287-327."""
    out = spec * 1
    xb, xg = np.nonzero(bad)[0], np.nonzero(~bad)[0]
    if len(xg) == 0:
        out[~np.isfinite(out)] = 1
        return out
    pos = np.searchsorted(xg, xb)
    le, re = pos == 0, pos == len(xg)
    mid = ~le & ~re
    l1, l2 = lam[xg[pos[mid] - 1]], lam[xg[pos[mid]]]
    s1, s2 = spec[xg[pos[mid] - 1]], spec[xg[pos[mid]]]
    l0 = lam[xb[mid]]
    out[xb[le]] = spec[xg[0]]
    out[xb[re]] = spec[xg[-1]]
    out[xb[mid]] = (-(l1 - l0) * s2 + (l2 - l0) * s1) / (l2 - l1)
    return out


def ccf_preprocess_data(lam, spec0, espec, conf, badmask=None, maxerr=10):
    """make_ccf.py:330-414."""
    logl = np.linspace(conf['logl0'], conf['logl1'], conf['npoints'])
    clam = np.exp(logl)
    es, sp = espec.copy(), spec0.copy()
    bad = np.zeros(len(es), dtype=bool) if badmask is None else badmask
    filt = scipy.signal.medfilt(sp, 11)
    mede = np.nanmedian(es)
    if conf['continuum']:
        bad = bad | (es > maxerr * mede) | (filt <= 0)
    es[bad] = 1e9 * mede
    sp = fill_masked(lam, sp, bad)
    cont = fit_continuum(lam, sp, es, conf) if conf['continuum'] else 1
    ivar = 1. / es**2
    ivar[bad] = 0
    medv = np.median(sp)
    cont = np.maximum(1e-2 * medv, cont) if medv > 0 else np.maximum(cont, 1)
    cs = spec0 / cont
    ivar = cont**2 * ivar
    cs[bad] = 0
    xi = np.searchsorted(lam, clam) - 1
    ok = (xi >= 0) & (xi <= len(lam) - 2)
    r1, r2 = np.zeros(len(logl)), np.zeros(len(logl))
    li = xi[ok]
    ri = li + 1
    rw = (clam[ok] - lam[li]) / (lam[ri] - lam[li])
    lw = 1 - rw
    r1[ok] = lw * cs[li] + rw * cs[ri]
    liv, riv = ivar[li], ivar[ri]
    r2[ok] = liv * riv / (lw**2 * riv + rw**2 * liv + ((liv * riv) == 0).astype(int))
    return r1, r2


def ccf_preprocess_model(logl, lam_m, model, vsini, conf):
    """make_ccf.py:167-212."""
    m = rotational_broaden(lam_m, model, vsini) if vsini else model
    if conf['continuum']:
        cont = fit_continuum(lam_m, m, np.maximum(m * 1e-5, 1e-2 * np.median(m)), conf)
        cont = np.maximum(cont, 1e-2 * np.median(cont))
    else:
        cont = 1
    return scipy.interpolate.interp1d(np.log(lam_m), m / cont, bounds_error=False,
                                      fill_value=1)(logl)


def build_ccf_bank(setup, conf, every=10, vsinis=(0.,)):
    """make_ccf.py:417-493 in memory: returns dict(fft, fft2, models, params,
    vsinis, parnames, ccfconf).  Node subsampling is a plain stride (the
    reference orders by a Morton key first, make_ccf.py:458-461; the choice of
    nodes is immaterial to the arithmetic under test)."""
    logl = np.linspace(conf['logl0'], conf['logl1'], conf['npoints'])
    inds = np.arange(0, setup['dats'].shape[0], every)
    vec = setup['vec'].T[inds].copy()
    vec[:, 0] = 10**vec[:, 0]
    models, params, vs = [], [], []
    for k, i in enumerate(inds):
        spec = np.exp(setup['dats'][i].astype(np.float64))
        for v in vsinis:
            models.append(ccf_preprocess_model(logl, setup['lam'], spec, v, conf))
            params.append(vec[k])
            vs.append(v)
    models = np.array(models)
    return dict(fft=np.fft.rfft(models, axis=1), fft2=np.fft.rfft(models**2, axis=1),
                models=models, params=np.array(params), vsinis=vs,
                parnames=setup['parnames'], ccfconf=conf)


def ccf_fit(specdata, config, banks, preprocessed=None):
    """fitter_ccf.py:62-253.  `banks` maps setup name -> build_ccf_bank()
    product; `preprocessed` optionally maps setup -> (proc_spec, proc_ivar) to
    bypass the host-side continuum fit."""
    maxvel = config.get('max_vel') or 1000
    nvg = 2 * int(maxvel * 1. / (config.get('vel_step0') or 2)) + 1
    vgrid = np.linspace(-maxvel, maxvel, nvg)
    if isinstance(specdata, SpecData):
        specdata = [specdata]
    states, sse, procs, steps = [], 0, {}, {}
    for sd in specdata:
        bank = banks[sd.name]
        conf = bank['ccfconf']
        if preprocessed is not None:
            ps, pi = preprocessed[sd.name]
        else:
            ps, pi = ccf_preprocess_data(sd.lam, sd.spec, sd.espec, conf, sd.badmask)
        procs[sd.name] = ps
        sse += (ps**2 * pi).sum()
        sconj = np.fft.rfft(ps * pi).conj()
        iconj = np.fft.rfft(pi).conj()
        step = (np.exp((conf['logl1'] - conf['logl0']) / conf['npoints']) - 1) * 3e5
        L = len(ps)
        off = L // 2
        vels = -((np.arange(L) + off) % L - off) * step
        sel = np.abs(vels) < (maxvel + step)
        assert sel.sum() % 2 == 1
        idx = np.roll(np.nonzero(sel)[0], sel.sum() // 2)[::-1]
        steps[sd.name] = step
        states.append(dict(sconj=sconj, iconj=iconj, bank=bank, idx=idx,
                           vels=vels[idx], cont=conf['continuum']))
    nfft = states[0]['bank']['fft'].shape[0]
    allc = np.zeros((nfft, nvg))
    for st in states:
        c0 = np.fft.irfft(st['bank']['fft'] * st['sconj'][None, :], axis=1)
        c1 = np.fft.irfft(st['bank']['fft2'] * st['iconj'][None, :], axis=1)
        chi = -2 * c0 + c1 if st['cont'] else -c0**2 / c1
        allc += scipy.interpolate.interp1d(st['vels'], chi[:, st['idx']], kind='linear',
                                           axis=1, assume_sorted=True)(vgrid)
    allc += sse
    bid = np.argmin(allc.min(axis=1))
    bccf = allc[bid]
    bpix = np.argmin(bccf)
    if bpix not in (0, len(bccf) - 1):
        co = np.polyfit(vgrid[bpix - 1:bpix + 2], bccf[bpix - 1:bpix + 2], deg=2)
        bvel = -co[1] / (2 * co[0]) if co[0] > 0 else vgrid[bpix]
    else:
        bvel = vgrid[bpix]
    if not np.isfinite(allc[bid, bpix]):
        raise RuntimeError('Cross-correlation step failed')
    b0 = states[0]['bank']
    bmodel = {sd.name: np.roll(banks[sd.name]['models'][bid],
                               int(bvel / steps[sd.name])) for sd in specdata}
    return dict(best_par=dict(zip(b0['parnames'], b0['params'][bid])), best_vel=bvel,
                best_ccf=bccf, best_vsini=b0['vsinis'][bid], best_model=bmodel,
                proc_spec=procs, vel_grid=vgrid, best_id=int(bid), all_chisqs=allc)
