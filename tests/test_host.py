"""Host-side logic of the product that needs no GPU: CCF preprocessing against
the reference's own outputs, the HDF5 dictionary decoder, optimiser helpers."""
import types

import numpy as np
import pytest

from helpers import close, unpack_objects
from rvspecfit_b200 import bank_io, make_ccf, vel_fit


def test_preprocess_data_matches_reference(golden):
    """make_ccf.preprocess_data (host, row f3) vs the reference's proc_spec /
    proc_ivar stored in tests/golden/ccf.npz."""
    g = golden('ccf')
    for tag, shapes in (('rvs', ('gaiarvs',)), ('two', ('desi_b', 'desi_r'))):
        for i, o in enumerate(unpack_objects(g, tag + '_')):
            for a, (nm, lam, sp, es, bad) in enumerate(o['arms']):
                c = g[f'{tag}_{nm}_conf']
                conf = make_ccf.get_ccf_config(c[0], c[1], int(c[2]))
                assert np.isclose(conf['splinestep'], c[3])
                ps, pi = make_ccf.preprocess_data(lam, sp, es, ccfconf=conf, badmask=bad)
                close(ps, g[f'{tag}_{i}_{a}_proc_spec'], rtol=1e-7, atol=1e-9, what='proc_spec')
                close(pi, g[f'{tag}_{i}_{a}_proc_ivar'], rtol=1e-6, atol=1e-12 * pi.max(),
                      what='proc_ivar')


def test_interp_masker_edges():
    lam = np.arange(10.)
    spec = np.arange(10.) * 2
    bad = np.zeros(10, dtype=bool)
    bad[[0, 1, 4, 5, 9]] = True
    out = make_ccf.interp_masker(lam, spec, bad)
    assert np.array_equal(out, [4, 4, 4, 6, 8, 10, 12, 14, 16, 16])
    allbad = make_ccf.interp_masker(lam, np.full(10, np.nan), np.ones(10, dtype=bool))
    assert np.array_equal(allbad, np.ones(10))


class _FakeItem:
    def __init__(self, val, kind):
        self.val, self.attrs = val, {'type': kind}
        self.dtype = np.asarray(val).dtype if not isinstance(val, bytes) else np.dtype('S')

    def __getitem__(self, key):
        return self.val if key == () else np.asarray(self.val)[key]


class _FakeGroup(dict):
    def __init__(self, d, kind=None):
        super().__init__(d)
        self.attrs = {} if kind is None else {'type': kind}


def test_h5_dictionary_decoder():
    """The typed-attribute scheme of the reference serializer
    (serializer.py:112-157), decoded without h5py through stand-in objects."""
    fake = types.SimpleNamespace(Group=_FakeGroup, Dataset=_FakeItem)
    root = _FakeGroup({
        'lam': _FakeItem(np.arange(4.), 'ndarray'),
        'parnames': _FakeItem(np.array(['teff', 'logg'], dtype=object), 'list'),
        'log_step': _FakeItem(np.bool_(True), 'scalar'),
        'revision': _FakeItem(b'v1', 'str'),
        'nothing': _FakeItem(0, 'None'),
        'uvecs': _FakeGroup({'__item_0': _FakeItem(np.array([1., 2.]), 'ndarray'),
                             '__item_1': _FakeItem(np.array([3., 4., 5.]), 'ndarray')},
                            'flattened_list'),
        'mapper_args': _FakeGroup({'__item_0': _FakeItem(np.array([0]), 'list')},
                                  'flattened_tuple'),
    })
    d = bank_io._decode(root, fake)
    assert d['parnames'] == ['teff', 'logg'] and d['revision'] == 'v1' and d['nothing'] is None
    assert isinstance(d['uvecs'], list) and len(d['uvecs'][1]) == 3
    assert isinstance(d['mapper_args'], tuple) and list(d['mapper_args'][0]) == [0]


def test_simplex_and_hessian_helpers():
    """Deterministic pieces of vel_fit that the batched driver shares with the
    single-object path."""
    vm = vel_fit.VSiniMapper(500)
    cur, simp = vel_fit._get_simplex_start(12., fixParam=[], specParamNames=('teff', 'logg'),
                                           paramDict0={'teff': 5000, 'logg': 2, 'vsini': 3},
                                           vsiniMapper=vm, fitVsini=True)
    assert simp.shape == (5, 4) and np.array_equal(simp[0], cur)
    R = np.random.RandomState(43434)
    assert np.allclose(simp[1:], cur + np.array([5, 3, 300, 0.5]) * R.normal(size=(4, 4)))
    H = vel_fit.central_hessian(lambda x: 0.5 * (3 * x[0]**2 + x[0] * x[1] + 2 * x[1]**2),
                                [0.3, -0.2], [1e-2, 1e-2])
    assert np.allclose(H, [[3, 0.5], [0.5, 2]], atol=1e-8)
    err, cov, bad = vel_fit._uncertainties_from_hessian(np.array([[4., 0.], [0., -1.]]))
    assert bad and err[0] == 0.5


def test_resolution_matrix_host_side(golden):
    """Host part of the resolution-matrix mode (no GPU): construct_resol_mat against the
    reference's matrices (tests/golden/resol.npz), the band-row layout of rvs_obs.d_resol,
    matrices stored with descending offsets (desi_fit.py:746), the resol_params views."""
    import pytest
    import scipy.sparse
    from rvspecfit_b200 import spec_fit
    g, gr = golden('chisq'), golden('resol')
    objs = unpack_objects(g, 'one_')
    for i, o in enumerate(objs):
        lam = o['arms'][0][1]
        rm = spec_fit.construct_resol_mat(lam, resol=float(gr['one_R'][i]))
        dia = scipy.sparse.dia_matrix(rm.mat)
        assert np.array_equal(dia.offsets, gr[f'one_{i}_offsets'])
        assert np.isclose(dia.data.sum(), gr[f'one_{i}_data_sum'], rtol=1e-13)
        if i == 0:
            close(dia.data, gr['one_0_data'], rtol=1e-13)
        # band rows by OUTPUT pixel reproduce the matrix product
        n = len(lam)
        offs, rows = spec_fit._band_rows(rm.mat, n)
        assert offs.dtype == np.int32 and np.all(np.diff(offs) > 0) and 0 in offs
        x = np.random.RandomState(i).normal(size=n)
        y = np.zeros(n)
        for k, off in enumerate(offs):
            p = np.arange(max(0, -off), min(n, n - off))
            y[p] += rows[k, p] * x[p + off]
        close(y, rm.mat @ x, rtol=1e-12, atol=1e-14)
        close(spec_fit.convolve_resol(x, rm), rm.mat @ x, rtol=0, atol=0)
    # DESI's storage order: offsets w2 .. -w2
    nm, lam, sp, es, bad = unpack_objects(g, 'desi_')[0]['arms'][0]
    offs, data = gr['desi_0_offsets'], gr['desi_0_data']
    up = scipy.sparse.dia_matrix((data, offs), shape=(len(lam), len(lam)))
    down = scipy.sparse.dia_matrix((data[::-1], offs[::-1]), shape=(len(lam), len(lam)))
    a, b = spec_fit._band_rows(up, len(lam)), spec_fit._band_rows(down, len(lam))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and len(a[0]) == 11
    with pytest.raises(ValueError):
        spec_fit._band_rows(up, len(lam) - 1)
    # SpecData carries the matrix; resol_params makes cached views and refuses both
    rm = spec_fit.ResolMatrix(up)
    sd = spec_fit.SpecData(nm, lam, sp, es, badmask=bad)
    v1 = spec_fit._with_resol_params([sd], {nm: rm})
    v2 = spec_fit._with_resol_params([sd], {nm: rm})
    assert v1[0] is v2[0] and v1[0].resolution is rm and v1[0]._band is not None
    assert sd.resolution is None
    with pytest.raises(ValueError):
        spec_fit._with_resol_params(v1, {nm: rm})
    # spectra sharing one matrix object share its band rows
    sd2 = spec_fit.SpecData(nm, lam, sp * 2, es, badmask=bad, resolution=rm)
    assert sd2._band is v1[0]._band


def test_preprocess_many_pool_equals_serial(golden):
    """The pool of host processes returns exactly what the serial loop does, in order."""
    g = golden('ccf')
    jobs = []
    for tag in ('rvs', 'two'):
        for o in unpack_objects(g, tag + '_'):
            for nm, lam, sp, es, bad in o['arms']:
                c = g[f'{tag}_{nm}_conf']
                jobs.append((lam, sp, es, bad, make_ccf.get_ccf_config(c[0], c[1], int(c[2]))))
    jobs = (jobs * 3)[:8]
    serial = make_ccf.preprocess_many(jobs, workers=1)
    pooled = make_ccf.preprocess_many(jobs, workers=2)
    again = make_ccf.preprocess_many(jobs[::-1], workers=2)[::-1]     # persistent pool
    make_ccf._close_pool()
    for a, b, c in zip(serial, pooled, again):
        for k in range(2):
            assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k])


def test_get_specdata_matches_reference(golden):
    """desi_data.get_specdata against the reference's desi_fit.get_specdata
    (desi/desi_fit.py:781-888) on synthetic frames with zero / negative / non-finite
    inverse variances, masked runs, clamped errors, a dropped arm and a zero-median arm,
    with and without resolution matrices (fixture specdata.npz)."""
    import scipy.sparse
    from rvspecfit_b200 import desi_data
    g = golden('specdata')
    setups = ['b', 'r', 'z']
    waves = {s: g[f'wave_{s}'] for s in setups}
    fluxes = {s: g[f'flux_{s}'] for s in setups}
    ivars = {s: g[f'ivar_{s}'] for s in setups}
    masks = {s: g[f'mask_{s}'] for s in setups}
    nfib = len(fluxes['b'])
    resol = {s: np.tile(g[f'resol_{s}'][None], (nfib, 1, 1)) for s in setups}
    sig0 = {'b': 0.5, 'r': 0.45, 'z': 0.4}
    ncmp = 0
    for mode, kw in (('plain', {}), ('resol', dict(use_resolution_matrix=True,
                                                      lsf_sigma0_angstrom=sig0))):
        for f in range(nfib):
            sds = desi_data.get_specdata(waves, fluxes, ivars, masks, resol, f, setups, **kw)
            want = [str(_) for _ in g[f'{mode}_{f}_names']]
            assert [sd.name for sd in sds or ()] == want, (mode, f)
            for sd in sds or ():
                assert np.array_equal(sd.spec, g[f'{mode}_{f}_{sd.name}_spec'], equal_nan=True)
                assert np.array_equal(sd.espec, g[f'{mode}_{f}_{sd.name}_espec'])
                assert np.array_equal(sd.badmask, g[f'{mode}_{f}_{sd.name}_bad'])
                key = f'{mode}_{f}_{sd.name}_resol_data'
                if key in g:
                    dia = scipy.sparse.dia_matrix(sd.resolution.mat)
                    assert np.array_equal(dia.offsets, g[f'{mode}_{f}_{sd.name}_resol_offsets'])
                    assert np.allclose(dia.data, g[key], rtol=1e-13, atol=1e-15)
                    ncmp += 1
    assert ncmp >= 4
    # arm 'r' of the last fibre is entirely masked -> dropped, as in the reference
    assert 'desi_r' not in [str(_) for _ in g[f'plain_{nfib - 1}_names']]


def test_template_library_change_clears_caches():
    """a config pointing at another template library must not be served
    the first library's banks (reference spec_inter.py:321-324)."""
    from rvspecfit_b200 import fitter_ccf, spec_fit, spec_inter
    ic = spec_inter.interp_cache
    saved = (dict(ic.interps), ic.template_lib, dict(fitter_ccf.CCFCache.banks),
             dict(spec_fit._engine_cache))
    try:
        ic.interps.clear()
        ic.interps['arm'] = 'interpolator of library A'
        ic.template_lib = 'libA/'
        fitter_ccf.CCFCache.banks['arm'] = 'ccf bank of library A'
        spec_fit._engine_cache['k'] = 'engine of library A'
        assert spec_inter.getInterpolator('arm', {'template_lib': 'libA/'}) == \
            'interpolator of library A'
        with pytest.raises(Exception):      # library B has no files: the old bank is NOT served
            spec_inter.getInterpolator('arm', {'template_lib': 'libB/'})
        assert 'arm' not in ic.interps and not fitter_ccf.CCFCache.banks
        assert not spec_fit._engine_cache
    finally:
        ic.interps.clear()
        ic.interps.update(saved[0])
        ic.template_lib = saved[1]
        fitter_ccf.CCFCache.banks.clear()
        fitter_ccf.CCFCache.banks.update(saved[2])
        spec_fit._engine_cache.clear()
        spec_fit._engine_cache.update(saved[3])
