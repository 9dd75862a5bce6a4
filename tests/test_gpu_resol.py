"""Resolution-matrix mode on the GPU (reference
spec_fit.py:410-492, 922-929, desi/desi_fit.py:723-748) against values the reference
itself produced (tests/golden/resol.npz) and against the oracle.  Needs a B200.

The three device routes are covered: the per-trial scan kernel (single evaluations, model
output), the tensor-core RV-scan kernel (several trials per template) and the fused
optimiser-phase evaluation with its resol_apply stage (general and TMA gather)."""
import numpy as np
import pytest
import scipy.sparse

import oracle
from helpers import PROBE_PARAMS, close, config, relerr, setup, unpack_objects
from rvspecfit_b200 import batch_fit, spec_fit, spec_inter, vel_fit

pytestmark = pytest.mark.gpu
CHI_RTOL = 1e-9


def _register(st, name=None):
    bank = spec_inter.bank_from_setup(st, kind='regulargrid', name=name)
    spec_inter.register_bank(bank, template_lib='synthetic/')
    return bank


def _dia(offsets, data, n, cls):
    return cls(scipy.sparse.dia_matrix((data, offsets), shape=(n, n)))


def test_construct_resol_mat_matches_reference(golden):
    g, gr = golden('chisq'), golden('resol')
    for i, o in enumerate(unpack_objects(g, 'one_')):
        lam = o['arms'][0][1]
        dia = scipy.sparse.dia_matrix(
            spec_fit.construct_resol_mat(lam, resol=float(gr['one_R'][i])).mat)
        assert np.array_equal(dia.offsets, gr[f'one_{i}_offsets'])
        assert np.isclose(dia.data.sum(), gr[f'one_{i}_data_sum'], rtol=1e-13)
        if i == 0:
            close(dia.data, gr['one_0_data'], rtol=1e-13)
    x = np.random.RandomState(3).normal(size=len(lam))
    rm = spec_fit.construct_resol_mat(lam, width=1.3)
    close(spec_fit.convolve_resol(x, rm), rm.mat.toarray() @ x, rtol=1e-12, atol=1e-14)


def test_get_chisq_with_resolution_matches_reference(golden):
    g, gr = golden('chisq'), golden('resol')
    _register(setup('test', 'tiny', 3, name='test'))
    objs = unpack_objects(g, 'one_')
    cfg, ev, opts = config(), g['one_eval'], {'npoly': 15}
    K = len(ev)
    on = np.arange(K) < K - 2       # the two off-grid points carry a float32 exp
    for i, o in enumerate(objs):
        nm, lam, sp, es, bad = o['arms'][0]
        rm = spec_fit.construct_resol_mat(lam, resol=float(gr['one_R'][i]))
        sd_res = [spec_fit.SpecData('test', lam, sp, es, badmask=bad, resolution=rm)]
        sd_plain = [spec_fit.SpecData('test', lam, sp, es, badmask=bad)]
        rots = [None if e[5] < 0 else (e[5],) for e in ev]
        # single evaluations through the reference-shaped call (fused path + resol_apply)
        got = np.array([spec_fit.get_chisq(sd_res, e[0], tuple(e[1:5]), r, options=opts,
                                           config=cfg) for e, r in zip(ev, rots)])
        assert relerr(got[on], gr['one_chisq'][i][on]) < CHI_RTOL
        assert relerr(got, gr['one_chisq'][i]) < 1e-6
        got = np.array([spec_fit.get_chisq(sd_plain, e[0], tuple(e[1:5]), r, options=opts,
                                           config=cfg, resol_params={'test': rm})
                        for e, r in zip(ev, rots)])
        assert relerr(got[on], gr['one_chisq_resol_params'][i][on]) < CHI_RTOL
        with pytest.raises(ValueError):
            spec_fit.get_chisq(sd_res, 0., tuple(ev[0][1:5]), None, options=opts, config=cfg,
                               resol_params={'test': rm})
        # the matrix together with the other switches of row a18: nearest-knot lookup
        # (per-trial kernel), systematic error + no off-grid penalty (fused path)
        got = np.array([spec_fit.get_chisq(sd_res, e[0], tuple(e[1:5]), r, options=opts,
                                           config=cfg, fast_interp=True)
                        for e, r in zip(ev, rots)])
        assert relerr(got[on], gr['one_chisq_fast'][i][on]) < CHI_RTOL
        got = np.array([spec_fit.get_chisq(sd_res, e[0], tuple(e[1:5]), r, options=opts,
                                           config=cfg, espec_systematic=float(gr['one_sys'][i]),
                                           outside_penalty=False) for e, r in zip(ev, rots)])
        assert relerr(got[on], gr['one_chisq_sys_nopen'][i][on]) < CHI_RTOL
        # nearest-knot lookup in the staged GEMM scan kernel == single evaluations
        eng = spec_fit.LikelihoodEngine([sd_res], cfg, opts)
        vg9 = np.linspace(-300, 300, 9)
        many = eng.evaluate([0], vg9[None, :], ev[None, 1, 1:5], np.array([max(ev[1, 5], 0.0)]),
                            fast_interp=True)[0]
        one = [spec_fit.get_chisq(sd_res, v, tuple(ev[1, 1:5]), (max(ev[1, 5], 0.0),),
                                  options=opts, config=cfg, fast_interp=True) for v in vg9]
        assert relerr(many, one) < CHI_RTOL
        # batched, fused and general path
        for fused in (True, False):
            eng = spec_fit.LikelihoodEngine([sd_res], cfg, opts, fused=fused)
            got = eng.evaluate(np.zeros(K, dtype=int), ev[:, 0], ev[:, 1:5],
                               np.where(ev[:, 5] < 0, 0.0, ev[:, 5]))
            norot = ev[:, 5] < 0        # the fixture has vsini=None there, not 0: same thing
            assert relerr(got[on], gr['one_chisq'][i][on]) < CHI_RTOL, fused
            assert norot.any()
        # model output (per-trial scan kernel)
        full = spec_fit.get_chisq(sd_res, ev[0][0], tuple(ev[0][1:5]), rots[0], options=opts,
                                  config=cfg, full_output=True)
        assert relerr(full['chisq'], gr[f'one_{i}_full_chisq']) < CHI_RTOL
        close(full['raw_models'][0], gr[f'one_{i}_full_raw'], rtol=1e-11, atol=1e-14)
        close(full['models'][0], gr[f'one_{i}_full_model'], rtol=1e-6)
        close(full['chisq_array'], gr[f'one_{i}_full_chisq_array'], rtol=1e-7)
        # RV scan (tensor-core scan kernel) + statistics
        fb = spec_fit.find_best(sd_res, np.arange(-400, 400, 10.),
                                [tuple(o['params']), PROBE_PARAMS[1]], rot_params=(25.,),
                                options=opts, config=cfg)
        for k in ('best_chi', 'best_vel', 'vel_err', 'kurtosis', 'skewness'):
            assert np.isclose(fb[k], gr[f'one_{i}_fb_{k}'], rtol=1e-7, atol=1e-9), k
        close(fb['probs'], gr[f'one_{i}_fb_probs'], rtol=1e-6, atol=1e-12)
        close(fb['best_param'], gr[f'one_{i}_fb_best_param'], rtol=1e-12)
        # the continuum-only fit applies the matrix to its unit template
        # (spec_fit.py:765-767; pinned by branches.npz in test_gpu_branches.py): the rows of
        # a truncated Gaussian matrix sum to 1 only to rounding, so the two values agree
        # closely but need not be identical
        a = spec_fit.get_chisq_continuum(sd_res, options=opts)['chisq_array']
        b = spec_fit.get_chisq_continuum(sd_plain, options=opts)['chisq_array']
        close(a, b, rtol=1e-6)


def test_three_arms_with_banded_matrices_match_reference(golden):
    """DESI-shaped object, one 11-diagonal matrix per arm (stored with descending
    offsets as desi_fit.py:746 does), fused path with the TMA gather, general path,
    and an RV scan."""
    g, gr = golden('chisq'), golden('resol')
    for k, a in enumerate(('desi_b', 'desi_r', 'desi_z')):
        _register(setup(a, 'tiny', 21 + k))
    o = unpack_objects(g, 'desi_')[0]
    sds = []
    for a, (nm, lam, sp, es, bad) in enumerate(o['arms']):
        offs, data = gr[f'desi_{a}_offsets'], gr[f'desi_{a}_data']
        rm = _dia(offs[::-1], data[::-1], len(lam), spec_fit.ResolMatrix)
        sds.append(spec_fit.SpecData(nm, lam, sp, es, badmask=bad, resolution=rm))
    cfg = config(min_vel=-1500, max_vel=1500)
    ev = g['desi_eval']
    for fused in (True, False):
        eng = spec_fit.LikelihoodEngine([sds], cfg, {'npoly': 10}, fused=fused)
        got = eng.evaluate(np.zeros(len(ev), dtype=int), ev[:, 0], ev[:, 1:5],
                           np.where(ev[:, 5] < 0, 0.0, ev[:, 5]))
        assert relerr(got, gr['desi_chisq']) < CHI_RTOL, fused
    vg = np.arange(-1500, 1500, 5.)[::6]
    chi = eng.evaluate([0], vg[None, :], np.array([o['params']]), np.array([12.]))
    assert relerr(chi[0], gr['desi_scan_chisq']) < CHI_RTOL


def test_mixed_batch_objects_with_and_without_matrix(golden):
    """One launch over objects with different matrices and one without (identity):
    each item equals its own single-object evaluation and the oracle."""
    g, gr = golden('chisq'), golden('resol')
    st = setup('test', 'tiny', 3, name='test')
    _register(st)
    oracle.register_setup(st)
    objs = unpack_objects(g, 'one_')
    cfg, ev, opts = config(), g['one_eval'][:6], {'npoly': 15}
    sds, osd = [], []
    for i, o in enumerate(objs + objs[:1]):
        nm, lam, sp, es, bad = o['arms'][0]
        kw = [dict(resol=1500.), dict(width=0.7), None][i]
        rm = None if kw is None else spec_fit.construct_resol_mat(lam, **kw)
        orm = None if kw is None else oracle.construct_resol_mat(lam, **kw)
        sds.append([spec_fit.SpecData('test', lam, sp, es, badmask=bad, resolution=rm)])
        osd.append([oracle.SpecData('test', lam, sp, es, bad, resolution=orm)])
    eng = spec_fit.LikelihoodEngine(sds, cfg, opts)
    obj = np.repeat(np.arange(3), len(ev))
    par = np.tile(ev, (3, 1))
    got = eng.evaluate(obj, par[:, 0], par[:, 1:5], np.where(par[:, 5] < 0, 0.0, par[:, 5]))
    want = np.array([oracle.get_chisq(osd[j], e[0], tuple(e[1:5]),
                                      None if e[5] < 0 else (e[5],), options=opts, config=cfg)
                     for j, e in zip(obj, par)])
    assert relerr(got, want) < CHI_RTOL
    # several trials per item: the GEMM scan kernel
    vg = np.linspace(-250, 250, 11)
    many = eng.evaluate(np.arange(3), np.tile(vg, (3, 1)), np.tile(ev[1, 1:5], (3, 1)),
                        np.full(3, 20.))
    one = np.array([[oracle.get_chisq(osd[j], v, tuple(ev[1, 1:5]), (20.,), options=opts,
                                      config=cfg) for v in vg] for j in range(3)])
    assert relerr(many, one) < CHI_RTOL


def test_process_with_resolution_matrix_matches_reference(golden):
    """A complete fit (vel_fit.process and batch_fit.process_batch) with a resolution
    matrix against the reference's own fit."""
    g, gr = golden('process'), golden('resol')
    _register(setup('test', 'test', 3, name='test'))
    nm, lam, sp, es, bad = unpack_objects(g, 'c1_')[0]['arms'][0]
    rm = spec_fit.construct_resol_mat(lam, resol=float(gr['proc_R']))
    sd = [spec_fit.SpecData(nm, lam, sp, es, badmask=bad, resolution=rm)]
    start = {'logg': 2, 'teff': 5000, 'feh': -0.2, 'alpha': 0.2, 'vsini': 0.1}
    one = vel_fit.process(sd, dict(start), fixParam=[], config=config(), options={'npoly': 15})
    two = batch_fit.process_batch([sd], [dict(start)], fixParam=[], config=config(),
                                  options={'npoly': 15})[0]
    for res in (one, two):
        par = np.array([res['param'][k] for k in ('teff', 'logg', 'feh', 'alpha')])
        assert abs(res['vel'] - gr['proc_vel']) < 0.01
        assert np.all(np.abs(par - gr['proc_param']) < 0.01 * gr['proc_param_err'])
        assert abs(res['chisq'] - gr['proc_chisq']) < 1e-6 * abs(res['chisq'])
        assert np.isclose(res['vel_err'], gr['proc_vel_err'], rtol=1e-3)
        close(res['yfit'][0], gr['proc_yfit'], rtol=1e-5)
    # resolParams dictionary instead of SpecData.resolution
    plain = [spec_fit.SpecData(nm, lam, sp, es, badmask=bad)]
    res = vel_fit.process(plain, dict(start), fixParam=[], config=config(),
                          options={'npoly': 15}, resolParams={nm: rm})
    assert abs(res['vel'] - gr['proc_vel']) < 0.01
    assert abs(res['chisq'] - gr['proc_chisq']) < 1e-6 * abs(res['chisq'])


def test_wide_band_takes_the_unstaged_scan_kernel(golden):
    """A matrix with more than RS_HW = 10 diagonals on a side (low resolving power) is
    applied by the scan kernel without the shared-memory staging: same values."""
    g = golden('chisq')
    st = setup('test', 'tiny', 3, name='test')
    _register(st)
    oracle.register_setup(st)
    nm, lam, sp, es, bad = unpack_objects(g, 'one_')[0]['arms'][0]
    rm = spec_fit.construct_resol_mat(lam, resol=500.)
    assert np.abs(scipy.sparse.dia_matrix(rm.mat).offsets).max() > 10
    sd = [spec_fit.SpecData('test', lam, sp, es, badmask=bad, resolution=rm)]
    osd = [oracle.SpecData('test', lam, sp, es, bad,
                           resolution=oracle.construct_resol_mat(lam, resol=500.))]
    cfg, opts, e = config(), {'npoly': 15}, g['one_eval'][1]
    eng = spec_fit.LikelihoodEngine([sd], cfg, opts)
    vg = np.linspace(-250, 250, 21)
    many = eng.evaluate([0], vg[None, :], e[None, 1:5], np.array([15.]))[0]
    want = [oracle.get_chisq(osd, v, tuple(e[1:5]), (15.,), options=opts, config=cfg) for v in vg]
    assert relerr(many, want) < CHI_RTOL
    one = eng.evaluate(np.zeros(len(vg), dtype=int), vg, np.tile(e[1:5], (len(vg), 1)),
                       np.full(len(vg), 15.))
    assert relerr(one, want) < CHI_RTOL
