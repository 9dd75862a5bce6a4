"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the
fixtures the reference itself produced (tests/golden).  Needs a B200.

Tolerances: BASELINE.json asks chi-square within 1e-6 relative, RV within
0.01 km/s, parameters within 1 % of sigma.  The kernels are fp64 end to end
and are held to 1e-9 relative on chi-square here (summation order and the
fp64 library functions differ from numpy's, nothing else)."""
import ctypes

import numpy as np
import pytest

import oracle
from helpers import PROBE_PARAMS, close, config, relerr, setup, unpack_objects
from rvspecfit_b200 import _cabi, _dev, spec_fit, spec_inter, vel_fit

pytestmark = pytest.mark.gpu
CHI_RTOL = 1e-9


def _register(st, kind='regulargrid', name=None):
    bank = spec_inter.bank_from_setup(st, kind=kind, name=name)
    spec_inter.register_bank(bank, template_lib='synthetic/')
    return bank


def _sd(obj, rename=None, cls=spec_fit.SpecData):
    return [cls(rename or nm, lam, sp, es, bad) for nm, lam, sp, es, bad in obj['arms']]


def test_spline_cabi_dropin(golden):
    """rvs_spline_construct / rvs_spline_eval against the reference's own
    outputs (tests/test_spline.py shapes)."""
    g = golden('kat')
    L = _cabi.lib()
    vp = _dev.hptr
    for tag in ('lin', 'log'):
        x, y, ex = [np.ascontiguousarray(g[f'spl_{tag}_{k}']) for k in ('x', 'y', 'ex')]
        n = len(x)
        A, B, C, D, h = [np.zeros(n - 1) for _ in range(5)]
        L.rvs_spline_construct(vp(x), vp(y), n, vp(A), vp(B), vp(C), vp(D), vp(h))
        for k, arr in zip('ABCD', (A, B, C, D)):
            close(arr, g[f'spl_{tag}_{k}'], rtol=1e-11, atol=1e-14), k
        out = np.zeros(len(ex))
        st = L.rvs_spline_eval(vp(ex), len(ex), n, vp(x), vp(h), vp(A), vp(B), vp(C), vp(D),
                               int(tag == 'log'), vp(out))
        assert st == 0
        close(out, g[f'spl_{tag}_val'], rtol=1e-11, atol=1e-12)
        bad = np.ascontiguousarray([x[0] - 1, x[5]])
        assert L.rvs_spline_eval(vp(bad), 2, n, vp(x), vp(h), vp(A), vp(B), vp(C), vp(D),
                                 int(tag == 'log'), vp(out)) == -1
        xb = x.copy()
        xb[1] += 0.3 * (x[1] - x[0])
        assert L.rvs_spline_eval(vp(ex), len(ex), n, vp(xb), vp(h), vp(A), vp(B), vp(C), vp(D),
                                 int(tag == 'log'), vp(out)) == -2


def test_template_interpolation(golden):
    g = golden('interp')
    pp = np.array(PROBE_PARAMS)
    for tag, holes in (('grid', 0), ('holes', 3)):
        st = setup('test', 'tiny', 3, holes=holes, name='p_' + tag)
        bank = _register(st)
        spec, outside = bank.template(pp)
        on = g[f'{tag}_outside'] == 0
        close(spec[on], g[f'{tag}_spec'][on], rtol=1e-13)
        # off-grid: the reference returns exp() of the float32 row evaluated in
        # float32 (numpy), reproduced as fp64 exp rounded to fp32: <= 1 fp32 ulp
        close(spec[~on], g[f"{tag}_spec"][~on], rtol=4e-7)
        assert (spec[~on] == g[f"{tag}_spec"][~on]).mean() > 0.5
        close(outside, g[f'{tag}_outside'], rtol=1e-13)
        it = spec_inter.getInterpolator('p_' + tag, config())
        close(it.eval(dict(zip(st['parnames'], pp[1]))), g[f'{tag}_spec'][1],
                           rtol=1e-13)
    st = setup('test', 'tiny', 3, name='p_tri')
    bank = _register(st, kind='triangulation')
    spec, outside = bank.template(pp)
    want_o = g['tri_outside']
    assert np.array_equal(np.isnan(outside), np.isnan(want_o))
    ok = ~np.isnan(want_o)
    close(outside[ok], want_o[ok], rtol=1e-9, atol=1e-12)
    close(spec[ok], g['tri_spec'][ok], rtol=1e-12)


def test_vsini_broadening_and_spline_vs_oracle(golden):
    """Broadened templates and their spline second derivatives against the
    oracle (scipy convolution + serial Thomas solve), incl. sub-pixel and very
    wide kernels."""
    g = golden('kat')
    lam, templ = g['conv_lam'], g['conv_templ']
    # a one-node bank whose single row is log(templ): interpolation = identity
    dats = np.tile(np.log(templ), (2, 1))
    bank = spec_inter.TemplateBank('conv', lam, dats, ('a',), kind='regulargrid',
                                   uvecs=[np.array([0., 1.])], idgrid=np.array([0, 1]),
                                   vecs=np.array([[0., 1.]]), log_ids=())
    for i, v in enumerate(g['conv_vsini']):
        ids = np.zeros((1, 2), dtype=np.int32)
        w = np.array([[1., 0.]])
        yz, st = bank.build(ids, w, np.array([v]))
        yz = _dev.download(yz)[0]
        close(yz[:, 0], g[f'conv_out{i}'], rtol=1e-12, atol=1e-14), v
        s = oracle.Spline(lam, np.ascontiguousarray(yz[:, 0]))
        z = np.concatenate([[0.], s.A * 6 * s.h])
        close(yz[:, 1], z, rtol=1e-9, atol=1e-12 * np.abs(z).max())


def test_basis_and_products():
    rs = np.random.RandomState(3)
    lam1 = np.linspace(4000, 5000, 777)
    lam2 = np.exp(np.linspace(np.log(6000), np.log(9000), 1234))
    sds = [spec_fit.SpecData('x', l, 1 + rs.uniform(size=len(l)), 0.1 + rs.uniform(size=len(l)))
           for l in (lam1, lam2, lam1)]
    b = spec_fit.SpectrumBatch(sds)
    assert list(b.grid_of) == [0, 1, 0]
    for npoly, rbf in ((1, True), (3, True), (10, True), (16, True), (7, False)):
        loglam, P, npp = b.basis(npoly, rbf)
        P = _dev.download(P)
        assert P.shape == (777 + 1234, npp) and npp % 2 == 0 and npp >= npoly
        close(P[:777, :npoly].T, oracle.continuum_basis(lam1, npoly, rbf), rtol=1e-13,
              atol=1e-15)
        close(P[777:, :npoly].T, oracle.continuum_basis(lam2, npoly, rbf), rtol=1e-12,
              atol=1e-14)
        assert (P[:, npoly:] == 0).all()
        close(_dev.download(loglam), np.log(np.concatenate([lam1, lam2])), rtol=1e-15)
    dn, einv, sumlog2 = [_dev.download(_) for _ in b.products(0.05)]
    es = np.sqrt(0.05**2 + b.h_espec**2)
    close(dn, b.h_spec / es, rtol=1e-15)
    close(einv, 1 / es, rtol=1e-15)
    close(sumlog2[1], 2 * np.log(es[777:777 + 1234]).sum(), rtol=1e-13)


@pytest.mark.parametrize('fused', [True, False])
def test_get_chisq_matches_reference(golden, fused):
    g = golden('chisq')
    st = setup('test', 'tiny', 3, name='test')
    _register(st)
    _register(setup('test', 'tiny', 3), kind='triangulation', name='test_tri')
    objs = unpack_objects(g, 'one_')
    ev = g['one_eval']
    cfg = config()
    for npoly, rbf in ((15, True), (5, True), (8, False)):
        opts = {'npoly': npoly, 'rbf_continuum': rbf}
        for name in ('test', 'test_tri'):
            want = g[f'one_chisq_{name}_{npoly}_{int(rbf)}']
            eng = spec_fit.LikelihoodEngine([_sd(o, name) for o in objs], cfg, opts,
                                            fused=fused)
            K = len(ev)
            for i in range(len(objs)):
                got = eng.evaluate(np.full(K, i), ev[:, 0], ev[:, 1:5],
                                   np.where(ev[:, 5] < 0, 0.0, ev[:, 5]))
                on = np.arange(K) < K - 2 if name == 'test' else np.ones(K, dtype=bool)
                assert relerr(got[on], want[i][on]) < CHI_RTOL, (name, npoly, rbf, i)
                # the two off-grid points carry the reference's float32 exp
                assert relerr(got, want[i]) < 1e-6, (name, npoly, rbf, i)
    # the reference-shaped single call
    sd = _sd(objs[0])
    e = ev[3]
    c = spec_fit.get_chisq(sd, e[0], tuple(e[1:5]), (e[5],), options={'npoly': 15}, config=cfg)
    assert abs(c - g['one_chisq_test_15_1'][0, 3]) < CHI_RTOL * abs(c)
    c = spec_fit.get_chisq(sd, ev[0, 0], tuple(ev[0, 1:5]), None, options={'npoly': 15},
                           config=cfg)
    assert abs(c - g['one_chisq_test_15_1'][0, 0]) < CHI_RTOL * abs(c)


def test_get_chisq_switches_match_reference(golden):
    """fast_interp (nearest-knot lookup in the scan kernels), espec_systematic as a
    scalar and as a per-setup dictionary, outside_penalty=False: reference values."""
    g, gs = golden('chisq'), golden('switches')
    _register(setup('test', 'tiny', 3, name='test'))
    objs = unpack_objects(g, 'one_')
    cfg, ev, opts = config(), g['one_eval'], {'npoly': 15}
    K = len(ev)
    for i, o in enumerate(objs):
        sd = _sd(o, 'test')
        sysv = float(gs[f'sys_{i}'])
        for key, kw in (('fast', dict(fast_interp=True)),
                        ('sys_scalar', dict(espec_systematic=sysv)),
                        ('sys_dict', dict(espec_systematic={'test': 2 * sysv})),
                        ('nopen', dict(outside_penalty=False))):
            got = np.array([spec_fit.get_chisq(sd, e[0], tuple(e[1:5]),
                                               None if e[5] < 0 else (e[5],), options=opts,
                                               config=cfg, **kw) for e in ev])
            on = np.arange(K) < K - 2       # the two off-grid points carry a float32 exp
            assert relerr(got[on], gs[key][i][on]) < CHI_RTOL, (key, i)
            assert relerr(got, gs[key][i]) < 1e-6, (key, i)
    # batched form, several trials per item: the GEMM scan kernel takes the same switch
    eng = spec_fit.LikelihoodEngine([_sd(objs[0], 'test')], cfg, opts)
    vg = np.linspace(-300, 300, 9)
    e = ev[1]
    many = eng.evaluate([0], vg[None, :], e[None, 1:5], np.array([max(e[5], 0.0)]),
                        fast_interp=True)[0]
    one = [spec_fit.get_chisq(_sd(objs[0], 'test'), v, tuple(e[1:5]), (max(e[5], 0.0),),
                              options=opts, config=cfg, fast_interp=True) for v in vg]
    assert relerr(many, one) < CHI_RTOL


def test_full_output_and_continuum(golden):
    g = golden('chisq')
    _register(setup('test', 'tiny', 3, name='test'))
    objs = unpack_objects(g, 'one_')
    sd = _sd(objs[0])
    ev = g['one_eval']
    fo = spec_fit.get_chisq(sd, ev[2, 0], tuple(ev[2, 1:5]), (ev[2, 5],),
                            options={'npoly': 15}, config=config(), full_output=True)
    assert abs(fo['chisq'] - g['one_full_chisq']) < 1e-8 * abs(fo['chisq'])
    close(fo['chisq_array'], g['one_full_chisq_array'], rtol=1e-7)
    assert np.array_equal(fo['npix_array'], g['one_full_npix'])
    close(fo['raw_models'][0], g['one_full_raw'], rtol=1e-12)
    close(fo['models'][0], g['one_full_model'], rtol=1e-6)
    cc = spec_fit.get_chisq_continuum(sd, options={'npoly': 15})['chisq_array']
    close(cc, g['one_cont'], rtol=1e-7)


def test_find_best_matches_reference(golden):
    g = golden('chisq')
    _register(setup('test', 'tiny', 3, name='test'))
    sd = _sd(unpack_objects(g, 'one_')[0])
    vg, plist = g['scan_vel_grid'], [tuple(_) for _ in g['scan_params']]
    cfg = config()
    eng = spec_fit.LikelihoodEngine([sd], cfg, {'npoly': 15})
    for tag, rot in (('norot', None), ('rot', (25.,))):
        vs = None if rot is None else np.full(len(plist), rot[0])
        chi = eng.evaluate(np.zeros(len(plist), dtype=int), np.tile(vg, (len(plist), 1)),
                           np.array(plist), vs)
        assert relerr(chi.T, g[f'scan_{tag}_chisq']) < CHI_RTOL
        fb = spec_fit.find_best(sd, vg, plist, rot_params=rot, options={'npoly': 15}, config=cfg)
        for k in ('best_chi', 'best_vel', 'vel_err', 'kurtosis', 'skewness'):
            assert np.isclose(fb[k], g[f'scan_{tag}_{k}'], rtol=1e-7, atol=1e-9), k
        assert abs(fb['best_vel'] - g[f'scan_{tag}_best_vel']) < 1e-6   # << 0.01 km/s
        close(fb['probs'], g[f'scan_{tag}_probs'], rtol=1e-6, atol=1e-12)
        close(fb['best_param'], g[f'scan_{tag}_best_param'], rtol=1e-12)


def test_desi_three_arm_matches_reference(golden):
    g = golden('chisq')
    for k, a in enumerate(('desi_b', 'desi_r', 'desi_z')):
        _register(setup(a, 'tiny', 21 + k))
    o = unpack_objects(g, 'desi_')[0]
    sd = _sd(o)
    cfg = config(min_vel=-1500, max_vel=1500)
    ev = g['desi_eval']
    for fused in (True, False):
        eng = spec_fit.LikelihoodEngine([sd], cfg, {'npoly': 10}, fused=fused)
        got = eng.evaluate(np.zeros(len(ev), dtype=int), ev[:, 0], ev[:, 1:5],
                           np.where(ev[:, 5] < 0, 0.0, ev[:, 5]))
        assert relerr(got, g['desi_chisq']) < CHI_RTOL
    vg = np.arange(-1500, 1500, 5.)
    chi = eng.evaluate([0], vg[None, :], np.array([o['params']]), np.array([12.]))
    assert relerr(chi[0], g['desi_scan_chisq']) < CHI_RTOL
    fb = spec_fit.find_best(sd, vg, [tuple(o['params'])], rot_params=(12.,),
                            options={'npoly': 10}, config=cfg)
    for k in ('best_chi', 'best_vel', 'vel_err', 'kurtosis', 'skewness'):
        assert np.isclose(fb[k], g[f'desi_scan_{k}'], rtol=1e-7, atol=1e-9), k


def test_ragged_arms_in_one_launch(golden):
    """Objects that lack some arms evaluated in the same launches as complete
    ones (absent arm = skipped item), against the oracle per object."""
    g = golden('chisq')
    for k, a in enumerate(('desi_b', 'desi_r', 'desi_z')):
        st = setup(a, 'tiny', 21 + k)
        _register(st)
        oracle.register_setup(st)
    o = unpack_objects(g, 'desi_')[0]
    sd, osd = _sd(o), _sd(o, cls=oracle.SpecData)
    cfg = config(min_vel=-1500, max_vel=1500)
    ev = g['desi_eval'][:6]
    subsets = [slice(0, 3), slice(0, 2), slice(2, 3), slice(1, 2)]
    eng = spec_fit.LikelihoodEngine([sd[s] for s in subsets], cfg, {'npoly': 10})
    K = len(ev)
    obj = np.tile(np.arange(len(subsets)), K)
    rep = np.repeat(np.arange(K), len(subsets))
    vs = np.where(ev[:, 5] < 0, 0.0, ev[:, 5])
    got = eng.evaluate(obj, ev[rep, 0], ev[rep, 1:5], vs[rep])
    for j, (i, e) in enumerate(zip(obj, rep)):
        want = oracle.get_chisq(osd[subsets[i]], ev[e, 0], tuple(ev[e, 1:5]), (vs[e],),
                                options={'npoly': 10}, config=cfg)
        assert abs(got[j] - want) < CHI_RTOL * abs(want), (i, e)


def test_tma_box_gather_identical_to_lane_gather(golden):
    """The copy-engine gather of the 2x2x2x2 corner box (rvs_gridbox) against the
    per-lane row gather: same rows, same accumulation order -> identical bits.
    Also a product whose nodes are listed in a scrambled order (the bank stores
    its rows in C order of the node table), interior / edge / off-grid points."""
    g = golden('chisq')
    st = setup('test', 'tiny', 3, name='test')
    objs = unpack_objects(g, 'one_')
    cfg = config()
    ev = g['one_eval']
    K = len(ev)
    pp = np.array(PROBE_PARAMS)
    vel = np.concatenate([ev[:, 0], np.linspace(-300, 300, len(pp))])
    par = np.concatenate([ev[:, 1:5], pp])
    vs = np.concatenate([np.where(ev[:, 5] < 0, 0.0, ev[:, 5]), np.linspace(0, 80, len(pp))])
    res = {}
    rs = np.random.RandomState(5)
    perm = rs.permutation(st['dats'].shape[0])
    inv = np.argsort(perm)
    for mode in ('box', 'lanes', 'scrambled'):
        s2 = dict(st)
        if mode == 'scrambled':     # node i of the product = node perm[i] of the original
            s2['dats'] = st['dats'][perm]
            s2['vec'] = st['vec'][:, perm]
            s2['idgrid'] = inv[st['idgrid']]
        bank = spec_inter.bank_from_setup(s2)
        assert bank.box is not None
        if mode == 'lanes':
            bank.box = None
        spec_inter.register_bank(bank, template_lib='synthetic/')
        eng = spec_fit.LikelihoodEngine([_sd(o, 'test') for o in objs], cfg, {'npoly': 10})
        res[mode] = np.array([eng.evaluate(np.full(len(vel), i), vel, par, vs)
                              for i in range(len(objs))])
    assert np.array_equal(res['box'], res['lanes'])
    assert np.array_equal(res['box'], res['scrambled'])
    want = g['one_chisq_test_15_1']
    assert res['box'].shape[1] == K + len(pp) and want.shape[1] == K


def test_graph_replay_identical_to_direct_launches(golden):
    """The evaluation call is captured into a CUDA graph the second time a
    configuration is seen and replayed afterwards: every replay must return the
    bits of the direct launch sequence, also after scratch buffers have grown
    (which invalidates the captured pointers) and with several evaluations in flight."""
    g = golden('chisq')
    for k, a in enumerate(('desi_b', 'desi_r', 'desi_z')):
        _register(setup(a, 'tiny', 21 + k))
    o = unpack_objects(g, 'desi_')
    cfg = config(min_vel=-1500, max_vel=1500)
    ev = g['desi_eval']
    vs = np.where(ev[:, 5] < 0, 0.0, ev[:, 5])
    sds = [_sd(x) for x in o]

    def run(eng, reps):
        obj = np.tile(np.arange(len(sds)), reps * len(ev))[:reps * len(ev)]
        ix = np.tile(np.arange(len(ev)), reps)
        return eng.evaluate(obj, ev[ix, 0], ev[ix, 1:5], vs[ix])

    direct = spec_fit.LikelihoodEngine(sds, cfg, {'npoly': 10})
    direct.use_graphs = False
    eng = spec_fit.LikelihoodEngine(sds, cfg, {'npoly': 10})
    assert eng.use_graphs
    want1, want4 = run(direct, 1), run(direct, 4)
    for _ in range(5 * eng.NSLOT):   # per in-flight slot: direct (x2: buffers of the other slots appear), capture + replay, replays
        assert np.array_equal(run(eng, 1), want1)
    assert eng.graph_kernel_launches > 0
    for _ in range(3 * eng.NSLOT):   # larger call: buffers grow, the old graphs are dropped
        assert np.array_equal(run(eng, 4), want4)
    for _ in range(3 * eng.NSLOT):
        assert np.array_equal(run(eng, 1), want1)
    # several evaluations in flight, alternating configurations
    obj1 = np.tile(np.arange(len(sds)), len(ev))[:len(ev)]
    for _ in range(3):
        hs = [eng.submit(obj1, ev[:, 0], ev[:, 1:5], vs) for _ in range(3)]
        for h in hs:
            assert np.array_equal(h.result(), want1)


def test_mixed_wavelength_grids_take_per_item_solve(golden):
    """Objects on different pixel grids in one engine: the continuum solve runs
    per item (gram_kernel) instead of as the shared-basis GEMM (gram_mma_kernel);
    both against the oracle, plus many trial points of ONE object (a group of
    the GEMM path made of a single object) and a batch that is not a multiple of
    the group size."""
    g = golden('chisq')
    st = setup('desi_b', 'tiny', 21)
    _register(st)
    oracle.register_setup(st)
    o = unpack_objects(g, 'desi_')[0]
    nm, lam, sp, es, bad = o['arms'][0]
    cuts = [slice(0, None), slice(100, 2500), slice(7, 1900)]
    sds = [[spec_fit.SpecData(nm, lam[c], sp[c], es[c], bad[c])] for c in cuts]
    osds = [[oracle.SpecData(nm, lam[c], sp[c], es[c], bad[c])] for c in cuts]
    cfg = config(min_vel=-1500, max_vel=1500)
    ev = g['desi_eval'][:7]
    vs = np.where(ev[:, 5] < 0, 0.0, ev[:, 5])
    K = len(ev)
    for npoly in (10, 13):
        opts = {'npoly': npoly}
        mixed = spec_fit.LikelihoodEngine(sds, cfg, opts)
        assert mixed.arms[nm]['batch'].obs(npoly, True).shared_grid == 0
        single = spec_fit.LikelihoodEngine(sds[1:2], cfg, opts)
        assert single.arms[nm]['batch'].obs(npoly, True).shared_grid == 1
        obj = np.repeat(np.arange(3), K)
        rep = np.tile(np.arange(K), 3)
        got = mixed.evaluate(obj, ev[rep, 0], ev[rep, 1:5], vs[rep])
        got1 = single.evaluate(np.zeros(K, dtype=int), ev[:, 0], ev[:, 1:5], vs)
        for j, (i, e) in enumerate(zip(obj, rep)):
            want = oracle.get_chisq(osds[i], ev[e, 0], tuple(ev[e, 1:5]), (vs[e],), options=opts,
                                    config=cfg)
            assert abs(got[j] - want) < CHI_RTOL * abs(want), (npoly, i, e)
            if i == 1:
                assert abs(got1[e] - want) < CHI_RTOL * abs(want), (npoly, e)


def test_locate_grid_bit_identical():
    """rvs_locate_grid reproduces the host vertex ids and weights bit for bit and
    flags exactly the points the host resolves through the KD-tree."""
    st = setup('test', 'tiny', 3, holes=3, name='test_holes')
    bank = _register(st)
    rs = np.random.RandomState(5)
    lo = np.array([st['uvecs'][i][0] for i in range(4)])
    hi = np.array([st['uvecs'][i][-1] for i in range(4)])
    q = lo + (hi - lo) * rs.uniform(-0.05, 1.05, size=(4000, 4))
    q[:50] = np.array([st['uvecs'][i][rs.randint(0, len(st['uvecs'][i]), 50)] for i in range(4)]).T
    q[50, 0] = np.nan
    q[51, 1] = np.inf
    params = q.copy()
    params[:, 0] = 10**q[:, 0]
    qm = spec_inter.map_params(params, bank.log_ids)
    ids, w, outside = bank.locate(params)
    K = len(q)
    d_q = _dev.upload(qm.T, np.float64)
    d_ids, d_w = _dev.empty((K, 16), np.int32), _dev.empty((K, 16), np.float64)
    d_flag = _dev.empty((K,), np.int32)
    rc = _cabi.lib().rvs_locate_grid(ctypes.byref(bank.gridmap), _dev.ptr(d_q), K, K,
                                     _dev.ptr(d_ids), _dev.ptr(d_w), _dev.ptr(d_flag), None,
                                     _dev.stream())
    assert rc == 0
    flag = _dev.download(d_flag).astype(bool)
    host_flag = (outside != 0) | (ids[:, 1] < 0)
    assert np.array_equal(flag, host_flag)
    assert 100 < flag.sum() < K - 100
    assert np.array_equal(_dev.download(d_ids)[~flag], ids[~flag])
    assert np.array_equal(_dev.download(d_w)[~flag], w[~flag])
    # with the off-grid resolution on the device: the KD-tree's node and distance
    d_out = _dev.empty((K,), np.float64)
    rc = _cabi.lib().rvs_locate_grid(ctypes.byref(bank.gridmap), _dev.ptr(d_q), K, K,
                                     _dev.ptr(d_ids), _dev.ptr(d_w), _dev.ptr(d_flag),
                                     _dev.ptr(d_out), _dev.stream())
    assert rc == 0
    flag2, out = _dev.download(d_flag).astype(bool), _dev.download(d_out)
    fin = np.isfinite(qm).all(axis=1)
    assert np.array_equal(flag2, ~fin) and np.isnan(out[~fin]).all()
    assert np.array_equal(_dev.download(d_ids)[fin], ids[fin])
    assert np.array_equal(_dev.download(d_w)[fin], w[fin])
    close(out[fin], outside[fin], rtol=1e-14, atol=0, what='off-grid measure')


def test_scan_stats_edge_cases():
    rs = np.random.RandomState(0)
    for nv, npar in ((7, 1), (600, 3), (33, 5)):
        v = np.sort(rs.uniform(-500, 500, nv))
        chi = rs.uniform(50, 60, size=(npar, nv))
        for where in ('mid', 'first', 'last'):
            c = chi.copy()
            j = {'mid': nv // 2, 'first': 0, 'last': nv - 1}[where]
            c[npar - 1, j] = 10. - (where == 'mid') * 0
            if where == 'mid':
                c[npar - 1, j - 1], c[npar - 1, j + 1] = 11., 12.
            out, pr = spec_fit.scan_stats(v[None], c[None])
            want = oracle.scan_statistics(v, c.T.copy())
            assert np.isclose(out[0, 0], want['best_chi'])
            assert np.isclose(out[0, 1], want['best_vel'], rtol=1e-10)
            assert np.isclose(out[0, 2], want['vel_err'], rtol=1e-10)
            assert np.isclose(out[0, 3], want['skewness'], rtol=1e-9, atol=1e-12)
            assert np.isclose(out[0, 4], want['kurtosis'], rtol=1e-9)
            assert int(out[0, 6]) == want['ibest']
            close(pr[0], want['probs'], rtol=1e-10)


def test_error_behaviour():
    st = setup('test', 'tiny', 3, name='test')
    _register(st)
    lam = np.linspace(4400, 5400, 500)      # bluer than the template coverage
    sd = [spec_fit.SpecData('test', lam, np.ones(500), np.full(500, 0.1))]
    with pytest.raises(RuntimeError):
        spec_fit.get_chisq(sd, 0., (5000., 2., -1., 0.2), None, options={'npoly': 5},
                           config=config())
    with pytest.raises(ValueError):
        spec_inter.getInterpolator('test', config()).eval({'teff': 5000.})


def test_process_matches_reference(golden):
    """BASELINE config 1 shape: vel_fit.process through the GPU likelihood vs the
    reference's own result.  RV within 0.01 km/s, parameters within 1 % of the
    reference's sigma (BASELINE.json)."""
    g = golden('process')
    _register(setup('test', 'test', 3, name='test'))
    objs = unpack_objects(g, 'c1_')
    start = {'logg': 2, 'teff': 5000, 'feh': -0.2, 'alpha': 0.2, 'vsini': 0.1}
    for i in (0, 1):
        sd = _sd(objs[i])
        for tag, cfg in (('nm', config(second_minimizer=False)), ('bfgs', config())):
            if i == 1 and tag == 'nm':
                continue
            res = vel_fit.process(sd, dict(start), fixParam=[], config=cfg,
                                  options={'npoly': 15})
            perr = g[f'c1_{i}_{tag}_param_err']
            par = np.array([res['param'][k] for k in ('teff', 'logg', 'feh', 'alpha')])
            assert abs(res['vel'] - g[f'c1_{i}_{tag}_vel']) < 0.01, (i, tag)
            assert np.all(np.abs(par - g[f'c1_{i}_{tag}_param']) < 0.01 * perr), (i, tag)
            assert abs(res['chisq'] - g[f'c1_{i}_{tag}_chisq']) < 1e-6 * abs(res['chisq'])
            assert np.isclose(res['vel_err'], g[f'c1_{i}_{tag}_vel_err'], rtol=1e-3)
            close(res['yfit'][0], g[f'c1_{i}_{tag}_yfit'], rtol=1e-5)
    res = vel_fit.process(_sd(objs[0]), dict(start), fixParam=['vsini', 'alpha'],
                          config=config(), options={'npoly': 15},
                          priors={'teff': (5200., 300.)})
    assert abs(res['vel'] - g['c1_0_fix_vel']) < 0.01
    assert abs(res['chisq'] - g['c1_0_fix_chisq']) < 1e-6 * abs(res['chisq'])
    fg = vel_fit.firstguess(_sd(objs[0]), config=config(), options={'npoly': 15},
                            paramsgrid={'logg': [1, 3, 4.5], 'teff': [4000, 6000, 9000],
                                        'feh': [-1.5, -0.5], 'alpha': [0.2]},
                            vsinigrid=(None, 50))
    got = np.array([fg[k] for k in ('teff', 'logg', 'feh', 'alpha')] + [fg.get('vsini', -1)])
    close(got, g['c1_fg'], rtol=1e-12)


def test_process_batch_matches_reference(golden):
    """batch_fit.process_batch (lock-step Nelder-Mead, pooled scipy BFGS, batched
    refinement scans and Hessian) against the reference's own fits of the same
    objects, all objects in one batch."""
    from rvspecfit_b200 import batch_fit
    g = golden('process')
    _register(setup('test', 'test', 3, name='test'))
    objs = unpack_objects(g, 'c1_')
    start = {'logg': 2, 'teff': 5000, 'feh': -0.2, 'alpha': 0.2, 'vsini': 0.1}
    sds = [_sd(objs[0]), _sd(objs[1]), _sd(objs[0])]
    for tag, cfg in (('bfgs', config()), ('nm', config(second_minimizer=False))):
        res = batch_fit.process_batch(sds, [dict(start) for _ in sds], fixParam=[], config=cfg,
                                      options={'npoly': 15})
        for j, i in enumerate((0, 1, 0)):
            if i == 1 and tag == 'nm':
                continue
            r = res[j]
            perr = g[f'c1_{i}_{tag}_param_err']
            par = np.array([r['param'][k] for k in ('teff', 'logg', 'feh', 'alpha')])
            assert abs(r['vel'] - g[f'c1_{i}_{tag}_vel']) < 0.01, (i, tag)
            assert np.all(np.abs(par - g[f'c1_{i}_{tag}_param']) < 0.01 * perr), (i, tag)
            assert abs(r['chisq'] - g[f'c1_{i}_{tag}_chisq']) < 1e-6 * abs(r['chisq'])
            assert np.isclose(r['vel_err'], g[f'c1_{i}_{tag}_vel_err'], rtol=1e-3)
            close(r['yfit'][0], g[f'c1_{i}_{tag}_yfit'], rtol=1e-5)
            assert r['minimize_success']
        # identical objects in one batch give identical fits
        assert res[0]['vel'] == res[2]['vel'] and res[0]['chisq'] == res[2]['chisq']
    # against the single-object driver, including uncertainties
    one = vel_fit.process(sds[1], dict(start), fixParam=[], config=config(second_minimizer=False),
                          options={'npoly': 15})
    b = res[1]
    assert abs(one['vel'] - b['vel']) < 1e-6 and abs(one['chisq'] - b['chisq']) < 1e-7 * abs(b['chisq'])
    for k in ('teff', 'logg', 'feh', 'alpha'):
        assert np.isclose(one['param'][k], b['param'][k], rtol=1e-7), k
        assert np.isclose(one['param_err'][k], b['param_err'][k], rtol=1e-3), k


def test_engine_reload_equals_fresh_engine(golden):
    """LikelihoodEngine.reload (new spectra of the same layout into the existing device
    buffers, captured graphs kept): identical bits to an engine built from scratch."""
    g = golden('chisq')
    _register(setup('test', 'tiny', 3, name='test'))
    objs = unpack_objects(g, 'one_')
    cfg, ev, opts = config(), g['one_eval'][:6], {'npoly': 15}
    a, b = _sd(objs[0], 'test'), _sd(objs[1], 'test')
    obj = np.repeat([0, 1], len(ev))
    par = np.tile(ev, (2, 1))
    vs = np.where(par[:, 5] < 0, 0.0, par[:, 5])
    eng = spec_fit.LikelihoodEngine([a, b], cfg, opts)
    for _ in range(3):          # captures the graphs
        first = eng.evaluate(obj, par[:, 0], par[:, 1:5], vs)
    eng.reload([b, a])
    swapped = eng.evaluate(obj, par[:, 0], par[:, 1:5], vs)
    fresh = spec_fit.LikelihoodEngine([b, a], cfg, opts).evaluate(obj, par[:, 0], par[:, 1:5], vs)
    assert np.array_equal(swapped, fresh)
    assert not np.array_equal(swapped, first)
    n = len(ev)
    assert np.array_equal(swapped[:n], first[n:]) and np.array_equal(swapped[n:], first[:n])
    # the scan and model-output paths see the new data too
    vg = np.linspace(-200, 200, 9)
    s1 = eng.evaluate([0], vg[None, :], ev[None, 1, 1:5], np.array([10.]))
    s2 = spec_fit.LikelihoodEngine([b], cfg, opts).evaluate([0], vg[None, :], ev[None, 1, 1:5],
                                                            np.array([10.]))
    assert np.array_equal(s1, s2)
    with pytest.raises(ValueError):
        eng.reload([a])
    short = [spec_fit.SpecData('test', a[0].lam[:-3], a[0].spec[:-3], a[0].espec[:-3])]
    with pytest.raises(ValueError):
        eng.reload([short, a])
