"""Seeded synthetic stellar library, instrument shapes and fake spectra.

There is no network in the build/bench environment, so PHOENIX-derived template
products cannot be used.  This module makes template grids and observed spectra
with the *shapes* of the instruments named in BASELINE.json:
the template wavelength grid follows the reference's preparation step
(reference make_interpol.py:313-323: log grid whose step equals `step` Å at the
middle of the range, padded by +-1000 km/s; spectra stored as float32 log-flux,
make_interpol.py:363-364), the node layout follows the regular-grid product
(reference make_nd.py:141-160).

The stellar model is a low-rank one,

    log F(theta, lam) = - sum_r c_r(theta) * L_r(lam),

with L_r >= 0 sums of Gaussian lines of a group r and c_r >= 0 smooth functions
of (teff, logg, feh, alpha).  A grid is therefore one small GEMM, which lets the
28600-node DESI-shaped grid be produced on the device in well under a second.
It is test/bench scaffolding: nothing here is on the timed path.
"""
import numpy as np

C_KMS = 299792.458
PARNAMES = ('teff', 'logg', 'feh', 'alpha')


def template_wavelengths(lam_lo, lam_hi, step, log_step=True, pad_kms=1000.):
    """Template wavelength grid as the reference preparation makes it
    (make_interpol.py:313-323)."""
    fac = 1 + pad_kms / C_KMS
    if not log_step:
        return np.arange(lam_lo / fac, (lam_hi + step) * fac, step)
    dln = np.log(1 + step / (0.5 * (lam_lo + lam_hi)))
    return np.exp(np.arange(np.log(lam_lo / fac), np.log(lam_hi * fac), dln))


# name -> template range/step, resolving power R(lam), observed grid maker
def _obs_linear(lo, hi, step):
    return lambda: np.arange(lo, hi + 0.5 * step, step)


SHAPES = {
    # reference tests/scripts/gen_test_templ_grid.sh + tests/test_fit_fake_grid.py
    'test': dict(t_lo=4550., t_hi=5450., t_step=1.0, resol=lambda x: 1000. + 0 * x,
                 obs=lambda: np.linspace(4600, 5400, 800)),
    # surveys/sdss/make_sdss.sh widened to BOSS
    'sdss': dict(t_lo=3500., t_hi=10500., t_step=1.0, resol=lambda x: 2000. + 0 * x,
                 obs=lambda: 10**(3.5563 + 1e-4 * np.arange(4600))),
    # surveys/desi/make_desi.sh:5-16
    'desi_b': dict(t_lo=3500., t_hi=5900., t_step=0.4, resol=lambda x: x / 1.55,
                   obs=_obs_linear(3600., 5800., 0.8)),
    'desi_r': dict(t_lo=5660., t_hi=7720., t_step=0.4, resol=lambda x: x / 1.55,
                   obs=_obs_linear(5760., 7620., 0.8)),
    'desi_z': dict(t_lo=7420., t_hi=9924., t_step=0.4, resol=lambda x: x / 1.8,
                   obs=_obs_linear(7520., 9824., 0.8)),
    # surveys/gaia_rvs/make_gaia.sh:5-8
    'gaiarvs': dict(t_lo=8410., t_hi=8750., t_step=0.05,
                    resol=lambda x: 11500. + 0 * x,
                    obs=lambda: np.arange(8460., 8700., 0.1)),
}

# node layouts; teff is interpolated in log10 (mapper log_ids=[0])
GRIDS = {
    'tiny': dict(teff=np.linspace(3500, 9000, 4), logg=np.linspace(0, 5, 3),
                 feh=np.linspace(-2, 0, 3), alpha=np.linspace(0, 1, 2)),
    'test': dict(teff=np.linspace(3000, 12000, 7), logg=np.linspace(0, 5, 7),
                 feh=np.linspace(-2, 0, 7), alpha=np.linspace(0, 1, 7)),
    'sdss': dict(teff=np.linspace(3000, 12000, 24), logg=np.linspace(0, 5, 9),
                 feh=np.linspace(-2.5, 0.5, 9), alpha=np.linspace(-0.2, 1.0, 4)),
    'desi': dict(teff=np.linspace(2500, 12000, 40), logg=np.linspace(0, 5, 11),
                 feh=np.linspace(-4, 0.5, 13), alpha=np.linspace(-0.2, 1.0, 5)),
    'small': dict(teff=np.linspace(3000, 12000, 10), logg=np.linspace(0, 5, 6),
                  feh=np.linspace(-2.5, 0.5, 6), alpha=np.linspace(0, 1.0, 4)),
}


class SynthLibrary:
    """Low-rank synthetic stellar library over a wavelength interval."""

    NGROUP = 6

    def __init__(self, lam_lo, lam_hi, seed=1, lines_per_angstrom=0.05):
        rs = np.random.RandomState(seed)
        span = lam_hi - lam_lo
        nl = max(12, int(span * lines_per_angstrom))
        self.cen = np.sort(rs.uniform(lam_lo - 0.02 * span, lam_hi + 0.02 * span, nl))
        self.group = rs.randint(0, 3, nl)  # groups 0..2 narrow metal-like lines
        self.strength = 10**rs.uniform(-1.6, -0.3, nl)
        self.width = np.full(nl, 0.08)
        # a few broad features: group 3 (hot, hydrogen-like), 4 (gravity
        # sensitive wings), 5 (cool-star bands)
        nb = max(3, int(span / 250.))
        for g, w, s in ((3, 6.0, 0.5), (4, 2.5, 0.25), (5, 25.0, 0.3)):
            c = rs.uniform(lam_lo, lam_hi, nb)
            self.cen = np.concatenate([self.cen, c])
            self.group = np.concatenate([self.group, np.full(nb, g)])
            self.strength = np.concatenate([self.strength, s * rs.uniform(0.5, 1.5, nb)])
            self.width = np.concatenate([self.width, np.full(nb, w)])

    def line_basis(self, lam, resol_func=None, xp=np):
        """L_r(lam), shape (NGROUP, npix); the LSF (sigma = lam/R/2.35) is folded
        in analytically, conserving each line's equivalent width."""
        lam = np.asarray(lam, dtype=np.float64)
        out = np.zeros((self.NGROUP, len(lam)))
        if resol_func is not None:
            sig_lsf = lam / resol_func(lam) / 2.35
        else:
            sig_lsf = np.zeros_like(lam)
        for k in range(len(self.cen)):
            w2 = self.width[k]**2 + sig_lsf**2
            reach = 8 * np.sqrt(w2.max())
            i0, i1 = np.searchsorted(lam, [self.cen[k] - reach, self.cen[k] + reach])
            if i1 <= i0:
                continue
            sl = slice(i0, i1)
            amp = self.strength[k] * self.width[k] / np.sqrt(w2[sl]) * \
                (1.0 if self.width[k] > 1 else 6.0)
            out[self.group[k], sl] += amp * np.exp(-0.5 * (lam[sl] - self.cen[k])**2 / w2[sl])
        return out

    @staticmethod
    def coeffs(teff, logg, feh, alpha):
        """c_r(theta) >= 0, shape (..., NGROUP)."""
        teff, logg, feh, alpha = [np.asarray(_, dtype=np.float64)
                                  for _ in (teff, logg, feh, alpha)]
        t = (teff - 3000.) / 9000.
        cool = np.exp(-2.2 * t)
        c0 = 10**(0.55 * feh) * cool * 1.6
        c1 = 10**(0.35 * feh) * (0.3 + cool) * (1 + 0.08 * (logg - 2.5))
        c2 = 10**(0.55 * (feh + alpha)) * cool * 1.6
        c3 = t**2 / (0.15 + t**2) * (1 + 0.06 * logg)
        c4 = 10**(0.22 * (logg - 2.5) + 0.25 * feh) * (0.4 + cool)
        c5 = np.exp(-7 * t) * 10**(0.3 * feh) * (1 + 0.1 * (logg - 2.5))
        return np.stack(np.broadcast_arrays(c0, c1, c2, c3, c4, c5), axis=-1)

    def logflux(self, basis, teff, logg, feh, alpha):
        return -(self.coeffs(teff, logg, feh, alpha) @ basis)


def make_nodes(layout):
    """Regular grid node table in the reference's order (files are sorted by
    teff, logg, feh, alpha: make_interpol.py:229-231).  Returns vec (4, Nnode)
    of physical parameters."""
    g = GRIDS[layout] if isinstance(layout, str) else layout
    mesh = np.meshgrid(g['teff'], g['logg'], g['feh'], g['alpha'], indexing='ij')
    return np.array([_.ravel() for _ in mesh])


def regular_index(vec_mapped):
    """uvecs / idgrid exactly as the regular-grid product stores them
    (make_nd.py:142-156)."""
    ndim = vec_mapped.shape[0]
    uu = [np.unique(vec_mapped[i], return_inverse=True) for i in range(ndim)]
    uvecs = [_[0] for _ in uu]
    idgrid = np.zeros([len(_) for _ in uvecs], dtype=int) - 1
    idgrid[tuple(_[1] for _ in uu)] = np.arange(vec_mapped.shape[1])
    return uvecs, idgrid


def make_setup(shape, layout, seed=1, holes=0, dats=None):
    """Build one spectral setup: dict(lam, dats float32 (Nnode,Npix), vec
    (mapped, log10 teff), uvecs, idgrid, parnames, log_step, lib, basis).
    `holes` removes that many interior nodes (idgrid == -1 there).  `dats` may
    supply the (memory-mapped) rows made by an earlier call with the same
    arguments."""
    sh = SHAPES[shape]
    lam = template_wavelengths(sh['t_lo'], sh['t_hi'], sh['t_step'])
    lib = SynthLibrary(sh['t_lo'], sh['t_hi'], seed=seed)
    basis = lib.line_basis(lam, sh['resol'])
    vec = make_nodes(layout)
    if holes:
        rs = np.random.RandomState(seed + 77)
        keep = np.ones(vec.shape[1], dtype=bool)
        keep[rs.choice(vec.shape[1], holes, replace=False)] = False
        vecfull = vec
        vec = vec[:, keep]
    if dats is None:
        dats = lib.logflux(basis, *vec).astype(np.float32)
    vmap = vec.copy()
    vmap[0] = np.log10(vmap[0])
    if holes:
        # keep every axis value present so that the index grid stays full-size
        vm_full = vecfull.copy()
        vm_full[0] = np.log10(vm_full[0])
        uvecs = [np.unique(vm_full[i]) for i in range(4)]
        idgrid = np.zeros([len(_) for _ in uvecs], dtype=int) - 1
        pos = tuple(np.searchsorted(uvecs[i], vmap[i]) for i in range(4))
        idgrid[pos] = np.arange(vmap.shape[1])
    else:
        uvecs, idgrid = regular_index(vmap)
    return dict(name=shape, lam=lam, dats=dats, vec=vmap, uvecs=uvecs,
                idgrid=idgrid, parnames=PARNAMES, log_step=True, lib=lib,
                resol=sh['resol'], shape=shape)


def doppler_factor(vel):
    """lam_rest = lam_obs * doppler_factor(v) (reference spec_fit.py:726-727)."""
    beta = vel / C_KMS
    return np.sqrt((1 - beta) / (1 + beta))


def fake_spectrum(setup, params, vel, sn, seed, lam=None, cont_amp=0.3,
                  bad_frac=0.0, scale=None):
    """One observed spectrum of a setup: the continuous library model at the
    Doppler-shifted wavelengths, times a smooth continuum, plus Gaussian noise.
    Returns lam, spec, espec, badmask."""
    rs = np.random.RandomState(seed)
    sh = SHAPES[setup['shape']]
    if lam is None:
        lam = sh['obs']()
    lam = np.asarray(lam, dtype=np.float64)
    lam_rest = lam * doppler_factor(vel)
    basis = setup['lib'].line_basis(lam_rest, setup['resol'])
    flux = np.exp(setup['lib'].logflux(basis, *params))
    x = (lam - lam[0]) / (lam[-1] - lam[0]) * 2 - 1
    cont = 1 + cont_amp * np.polynomial.chebyshev.chebval(x, rs.uniform(-1, 1, 4) / 2.)
    cont = np.maximum(cont, 0.2)
    if scale is None:
        scale = 10**rs.uniform(-1, 2)
    model = flux * cont * scale
    espec = np.abs(model) / sn + 1e-6 * scale
    spec = model + espec * rs.normal(size=len(lam))
    badmask = np.zeros(len(lam), dtype=bool)
    if bad_frac > 0:
        badmask = rs.uniform(size=len(lam)) < bad_frac
        espec = np.where(badmask, espec * 1000, espec)
    return lam, spec, espec, badmask


def random_params(layout, n, seed, margin=0.05):
    """n parameter vectors (teff, logg, feh, alpha) uniform inside a layout."""
    g = GRIDS[layout] if isinstance(layout, str) else layout
    rs = np.random.RandomState(seed)
    out = np.zeros((n, 4))
    for j, k in enumerate(PARNAMES):
        lo, hi = g[k][0], g[k][-1]
        if k == 'teff':
            lo, hi = max(lo, 3500.), min(hi, 9000.)
        d = (hi - lo) * margin
        out[:, j] = rs.uniform(lo + d, hi - d, n)
    return out


def fast_spectra(setup, params, vel, sn, seed, cont_amp=0.3, bad_frac=0.01):
    """Many observed spectra of one setup at once (bench / batch tests): the
    library's group basis is interpolated from the template grid to the
    Doppler-shifted wavelengths instead of being re-evaluated line by line.
    params (B,4), vel (B,), sn (B,).  Returns lam (npix,), spec, espec (B,npix),
    badmask (B,npix)."""
    rs = np.random.RandomState(seed)
    lam = SHAPES[setup['shape']]['obs']()
    B = len(vel)
    if 'basis' not in setup:
        setup['basis'] = setup['lib'].line_basis(setup['lam'], setup['resol'])
    coef = setup['lib'].coeffs(*np.asarray(params).T)          # (B, R)
    x = (lam - lam[0]) / (lam[-1] - lam[0]) * 2 - 1
    spec = np.empty((B, len(lam)))
    for i in range(B):
        lr = lam * doppler_factor(vel[i])
        lf = np.zeros(len(lam))
        for r in range(coef.shape[1]):
            lf -= coef[i, r] * np.interp(lr, setup['lam'], setup['basis'][r])
        cont = 1 + cont_amp * np.polynomial.chebyshev.chebval(x, rs.uniform(-1, 1, 4) / 2.)
        spec[i] = np.exp(lf) * np.maximum(cont, 0.2) * 10**rs.uniform(-1, 2)
    espec = np.abs(spec) / np.asarray(sn)[:, None] + 1e-6 * np.abs(spec).mean(axis=1)[:, None]
    spec = spec + espec * rs.normal(size=spec.shape)
    bad = rs.uniform(size=spec.shape) < bad_frac
    espec = np.where(bad, espec * 1000, espec)
    return lam, spec, espec, bad
