"""Generate tests/golden/*.npz by running the REFERENCE package itself.

Runs only in the build container (needs /root/reference).  It installs the
reference into the git-ignored baseline/_ref/ (baseline/install_ref.py: pip
install of a scratch copy, which builds its cffi spline as the reference's own
setup.py does; nothing enters this repository's history), stubs the absent
third-party imports (h5py, astropy, numdifftools, matplotlib --
baseline/ref_loader.py), injects seeded synthetic template
banks (rvspecfit_b200/synth.py) straight into the reference's caches
(spec_inter.interp_cache, fitter_ccf.CCFCache -- SURVEY.md §8c) and records what
the reference computes.  The resulting fixtures travel to the GPU box; the
reference does not.

    python tests/golden/make_golden.py            # all fixtures
    python tests/golden/make_golden.py kat chisq  # a subset

numdifftools is absent, so `process` fixtures are produced with
ndf.Hessian replaced by oracle.central_hessian: param_err / param_covar /
bad_hessian in those fixtures are NOT reference outputs ("parity unpinned").
"""
import os
import sys
import types

import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from rvspecfit_b200 import synth  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import PROBE_PARAMS  # noqa: E402

REF_SRC = '/root/reference/py/rvspecfit'


def load_reference():
    """The reference package as installed by baseline/install_ref.py (the same copy
    bench.py's reference arm times), third-party stubs from baseline/ref_loader.py."""
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import install_ref
    import ref_loader
    install_ref.install()
    return ref_loader.load(synth.PARNAMES)


def inject_grid(R, setup, name=None):
    import ref_loader
    return ref_loader.inject_grid(R, setup, name)


def inject_tri(R, setup, name):
    """Delaunay product assembled with the reference's own make_nd logic
    (make_nd.py:101-140) from in-memory arrays."""
    import scipy.spatial
    si = R.spec_inter
    vec = setup['vec'].astype(float)
    state = np.random.get_state()
    np.random.seed(1)
    vec = vec + np.random.uniform(-1e-6, 1e-6, size=vec.shape)
    np.random.set_state(state)
    edge = R.make_nd.getedgevertices(vec)
    near = scipy.spatial.cKDTree(vec.T).query(edge.T)[1]
    nspec = setup['dats'].shape[0]
    vec = np.hstack((vec, edge))
    specs = np.append(setup['dats'], np.array([setup['dats'][_] for _ in near]), axis=0)
    flags = np.concatenate((np.zeros(nspec), np.ones(edge.shape[1])))
    specs = specs.astype(np.float64)
    flags = flags.astype(np.float64)[:, None]
    tri = scipy.spatial.Delaunay(vec.astype(np.float64).T)
    si.interp_cache.template_lib = 'synthetic/'
    it = si.SpecInterpolator(name, si.TriInterp(tri, specs, exp=True),
                             si.TriInterp(tri, flags, exp=False), setup['lam'],
                             R.read_grid.LogParamMapper([0]), setup['parnames'],
                             log_step=True)
    si.interp_cache.interps[name] = it
    return it


CONFIG = dict(min_vel=-1000, max_vel=1000, vel_step0=5, max_vsini=500, min_vsini=0.1,
              min_vel_step=0.2, second_minimizer=True, template_lib='synthetic/')


def frozen_config(R, **kw):
    c = dict(CONFIG)
    c.update(kw)
    return R.utils.freezeDict(c)


def checksum(a):
    return float(np.asarray(a, dtype=np.float64).sum())


# ------------------------------------------------------------------ fixtures
def gold_kat(R):
    """Known-answer vectors of the native spline (tests/test_spline.py shapes),
    the vsini kernel / convolution and the continuum-marginalised chi-square."""
    rs = np.random.RandomState(11)
    out = {}
    for tag, x in (('lin', np.linspace(1000, 2000, 1000)),
                   ('log', 10**np.linspace(3, 4, 1000))):
        y = np.sin(x / 10) + rs.normal(size=len(x))
        ex = np.sort(rs.uniform(x[0], min(x[-1], 2000) - 1e-3, size=2000))
        s = R.spliner.Spline(x, y, log_step=(tag == 'log'))
        out.update({f'spl_{tag}_x': x, f'spl_{tag}_y': y, f'spl_{tag}_ex': ex,
                    f'spl_{tag}_A': s.A, f'spl_{tag}_B': s.B, f'spl_{tag}_C': s.C,
                    f'spl_{tag}_D': s.D, f'spl_{tag}_val': s(ex)})
    Rs = np.array([1e-3, 0.05, 0.4, 0.999, 1.0, 1.7, 3.2, 12.5, 40.01])
    out['vsini_R'] = Rs
    for i, r in enumerate(Rs):
        out[f'vsini_k{i}'] = R.spec_fit.compute_vsini_kernel(r)
    lam = synth.template_wavelengths(4550, 5450, 1.0)
    templ = 1 - 0.5 * np.exp(-0.5 * ((lam - 5000.3) / 1.5)**2) + 0.01 * rs.normal(size=len(lam))
    vs = np.array([0., 1e-7, 3., 30., 150., 499.])
    out['conv_lam'], out['conv_templ'], out['conv_vsini'] = lam, templ, vs
    for i, v in enumerate(vs):
        out[f'conv_out{i}'] = R.spec_fit.convolve_vsini(lam, templ, v)
    # chi-square kernel: cholesky route, svd route with coefficients
    npix = 700
    lam_o = np.linspace(4600, 5400, npix)
    t = 1 - 0.4 * np.exp(-0.5 * ((lam_o - 5003) / 2.)**2)
    es = 0.02 * (1 + 0.5 * rs.uniform(size=npix))
    sp = t * (1 + 0.2 * (lam_o - 5000) / 400) * 3.3 + es * rs.normal(size=npix)
    out.update(c0_lam=lam_o, c0_templ=t, c0_spec=sp, c0_espec=es)
    for npoly, rbf in ((5, True), (10, True), (15, True), (10, False), (2, True)):
        P = R.spec_fit.get_poly_basis(lam_o, npoly, rbf=rbf)
        tag = f'{npoly}_{int(rbf)}'
        out[f'c0_basis_{tag}'] = P
        out[f'c0_chol_{tag}'] = R.spec_fit.get_chisq0(sp, t, P, espec=es)
        c, co = R.spec_fit.get_chisq0(sp, t, P, get_coeffs=True, espec=es)
        out[f'c0_svd_{tag}'] = c
        out[f'c0_coeffs_{tag}'] = co
    np.savez_compressed(os.path.join(HERE, 'kat.npz'), **out)




def gold_interp(R):
    """Template evaluation: polylinear (with and without holes) and Delaunay."""
    out = {}
    full = synth.make_setup('test', 'tiny', seed=3)
    holed = synth.make_setup('test', 'tiny', seed=3, holes=3)
    out['params'] = np.array(PROBE_PARAMS)
    for tag, st in (('grid', full), ('holes', holed)):
        it = inject_grid(R, st, name='g_' + tag)
        out[f'{tag}_dats_sum'] = checksum(st['dats'])
        out[f'{tag}_spec'] = np.array([it.eval(p) for p in PROBE_PARAMS])
        out[f'{tag}_outside'] = np.array([float(it.outsideFlag(p)) for p in PROBE_PARAMS])
    it = inject_tri(R, full, 'g_tri')
    sp, of = [], []
    for p in PROBE_PARAMS:
        s = it.eval(p)
        sp.append(np.full(len(full['lam']), np.nan) if np.ndim(s) == 0 else s)
        of.append(float(it.outsideFlag(p)))
    out['tri_spec'], out['tri_outside'] = np.array(sp), np.array(of)
    np.savez_compressed(os.path.join(HERE, 'interp.npz'), **out)


def make_objects(arms, layout, n, seed0, sn_range=(20, 200), vel_sig=150., bad_frac=0.01,
                 lam_override=None):
    """n synthetic multi-arm objects; returns list of dict(params, vel, arms=[(name,lam,spec,espec,bad)])."""
    pars = synth.random_params(layout, n, seed0)
    rs = np.random.RandomState(seed0 + 1)
    objs = []
    for i in range(n):
        vel = rs.normal(0, vel_sig)
        sn = np.exp(rs.uniform(*np.log(sn_range)))
        data = []
        for a, st in enumerate(arms):
            lam = None if lam_override is None else lam_override[a]
            lam, spec, espec, bad = synth.fake_spectrum(st, pars[i], vel, sn,
                                                        seed0 + 100 * i + a, lam=lam,
                                                        bad_frac=bad_frac)
            data.append((st['name'], lam, spec, espec, bad))
        objs.append(dict(params=pars[i], vel=vel, sn=sn, arms=data))
    return objs


def specdata_of(R, obj):
    return [R.spec_fit.SpecData(nm, lam, sp, es, badmask=bad)
            for nm, lam, sp, es, bad in obj['arms']]


def pack_objects(objs, out, prefix):
    out[prefix + 'n'] = len(objs)
    out[prefix + 'params'] = np.array([o['params'] for o in objs])
    out[prefix + 'vel'] = np.array([o['vel'] for o in objs])
    for i, o in enumerate(objs):
        out[f'{prefix}{i}_names'] = np.array([a[0] for a in o['arms']])
        for a, (nm, lam, sp, es, bad) in enumerate(o['arms']):
            out[f'{prefix}{i}_{a}_lam'] = lam
            out[f'{prefix}{i}_{a}_spec'] = sp
            out[f'{prefix}{i}_{a}_espec'] = es
            out[f'{prefix}{i}_{a}_bad'] = bad


def gold_chisq(R):
    """get_chisq / find_best on a single-arm test-shape object (polylinear and
    Delaunay) and on a DESI-shaped 3-arm object."""
    out = {}
    cfg = frozen_config(R)
    st = synth.make_setup('test', 'tiny', seed=3)
    inject_grid(R, st, 'test')
    tri = dict(st)
    tri['name'] = 'test_tri'
    inject_tri(R, st, 'test_tri')
    objs = make_objects([st], 'tiny', 2, 500, bad_frac=0.02)
    pack_objects(objs, out, 'one_')
    out['one_dats_sum'] = checksum(st['dats'])
    rs = np.random.RandomState(5)
    # evaluation points: (vel, teff, logg, feh, alpha, vsini or -1 for None)
    ev = []
    for k in range(14):
        p = synth.random_params('tiny', 1, 900 + k)[0]
        ev.append([rs.uniform(-600, 600), *p, [-1, 0., 5., 40., 180.][k % 5]])
    ev.append([100., 9500., 2.0, -1.0, 0.2, 10.])     # off grid -> penalty
    ev.append([-50., 5000., 2.0, -2.6, 0.2, -1])      # off grid
    ev = np.array(ev)
    out['one_eval'] = ev
    for npoly, rbf in ((15, True), (5, True), (8, False)):
        opts = {'npoly': npoly, 'rbf_continuum': rbf}
        for name in ('test', 'test_tri'):
            res = np.zeros((len(objs), len(ev)))
            for i, o in enumerate(objs):
                sd = specdata_of(R, o)
                if name != 'test':
                    sd = [R.spec_fit.SpecData(name, s.lam, s.spec, s.espec, badmask=s.badmask)
                          for s in sd]
                for j, e in enumerate(ev):
                    rot = None if e[5] < 0 else (e[5],)
                    res[i, j] = R.spec_fit.get_chisq(sd, e[0], tuple(e[1:5]), rot,
                                                     options=opts, config=cfg)
            out[f'one_chisq_{name}_{npoly}_{int(rbf)}'] = res
    # full_output at one point
    sd = specdata_of(R, objs[0])
    fo = R.spec_fit.get_chisq(sd, ev[2, 0], tuple(ev[2, 1:5]), (ev[2, 5],),
                              options={'npoly': 15}, config=cfg, full_output=True)
    out['one_full_chisq'] = fo['chisq']
    out['one_full_chisq_array'] = np.array(fo['chisq_array'])
    out['one_full_npix'] = np.array(fo['npix_array'])
    out['one_full_model'] = fo['models'][0]
    out['one_full_raw'] = fo['raw_models'][0]
    out['one_cont'] = R.spec_fit.get_chisq_continuum(sd, options={'npoly': 15})['chisq_array']
    # RV scan
    vg = np.arange(-1000, 1000, 5.)
    plist = [tuple(objs[0]['params']), (5000., 2.0, -1.0, 0.2), (7000., 4.0, -0.5, 0.)]
    for tag, rot in (('norot', None), ('rot', (25.,))):
        chi = np.zeros((len(vg), len(plist)))
        cache = R.spec_fit.LRUDict(100)
        for j, p in enumerate(plist):
            for i, v in enumerate(vg):
                chi[i, j] = R.spec_fit.get_chisq(sd, v, p, rot, options={'npoly': 15},
                                                 config=cfg, cache=cache)
        fb = R.spec_fit.find_best(sd, vg, plist, rot_params=rot, options={'npoly': 15},
                                  config=cfg)
        out[f'scan_{tag}_chisq'] = chi
        for k in ('best_chi', 'best_vel', 'vel_err', 'kurtosis', 'skewness', 'probs'):
            out[f'scan_{tag}_{k}'] = fb[k]
        out[f'scan_{tag}_best_param'] = np.array(fb['best_param'])
    out['scan_vel_grid'] = vg
    out['scan_params'] = np.array(plist)

    # DESI-shaped three-arm object, tiny node layout, npoly 10, +-1500 km/s
    arms = [synth.make_setup(a, 'tiny', seed=21 + k)
            for k, a in enumerate(('desi_b', 'desi_r', 'desi_z'))]
    for a in arms:
        inject_grid(R, a)
    cfg3 = frozen_config(R, min_vel=-1500, max_vel=1500)
    o3 = make_objects(arms, 'tiny', 1, 700, bad_frac=0.01)
    pack_objects(o3, out, 'desi_')
    out['desi_dats_sum'] = np.array([checksum(a['dats']) for a in arms])
    sd3 = specdata_of(R, o3[0])
    ev3 = ev[:8].copy()
    ev3[:, 0] *= 2
    out['desi_eval'] = ev3
    out['desi_chisq'] = np.array([
        R.spec_fit.get_chisq(sd3, e[0], tuple(e[1:5]), None if e[5] < 0 else (e[5],),
                             options={'npoly': 10}, config=cfg3) for e in ev3])
    vg3 = np.arange(-1500, 1500, 5.)
    p3 = tuple(o3[0]['params'])
    cache = R.spec_fit.LRUDict(100)
    out['desi_scan_chisq'] = np.array([
        R.spec_fit.get_chisq(sd3, v, p3, (12.,), options={'npoly': 10}, config=cfg3,
                             cache=cache) for v in vg3])
    fb = R.spec_fit.find_best(sd3, vg3, [p3], rot_params=(12.,), options={'npoly': 10},
                              config=cfg3)
    for k in ('best_chi', 'best_vel', 'vel_err', 'kurtosis', 'skewness'):
        out[f'desi_scan_{k}'] = fb[k]
    np.savez_compressed(os.path.join(HERE, 'chisq.npz'), **out)


def gold_process(R):
    """vel_fit.process / firstguess on BASELINE config-1-shaped input (7^4
    regular grid, 800 px, npoly 15, vsini free) and on a DESI-shaped object."""
    out = {}
    cfg = frozen_config(R)
    st = synth.make_setup('test', 'test', seed=3)
    inject_grid(R, st, 'test')
    out['test_dats_sum'] = checksum(st['dats'])
    objs = make_objects([st], 'test', 3, 4000, sn_range=(60, 150), vel_sig=100.,
                        bad_frac=0.0)
    pack_objects(objs, out, 'c1_')
    start = {'logg': 2, 'teff': 5000, 'feh': -0.2, 'alpha': 0.2, 'vsini': 0.1}
    keys = ('vel', 'vel_err', 'vel_skewness', 'vel_kurtosis', 'vsini', 'chisq', 'logl')
    for i, o in enumerate(objs):
        sd = specdata_of(R, o)
        for tag, c in (('bfgs', cfg), ('nm', frozen_config(R, second_minimizer=False))):
            res = R.vel_fit.process(sd, dict(start), fixParam=[], config=c,
                                    options={'npoly': 15})
            for k in keys:
                out[f'c1_{i}_{tag}_{k}'] = res[k]
            out[f'c1_{i}_{tag}_param'] = np.array([res['param'][k] for k in synth.PARNAMES])
            out[f'c1_{i}_{tag}_param_err'] = np.array([res['param_err'][k]
                                                       for k in synth.PARNAMES])
            out[f'c1_{i}_{tag}_chisq_array'] = np.array(res['chisq_array'])
            out[f'c1_{i}_{tag}_yfit'] = res['yfit'][0]
            out[f'c1_{i}_{tag}_success'] = res['minimize_success']
        if i == 0:
            res = R.vel_fit.process(sd, dict(start), fixParam=['vsini', 'alpha'],
                                    config=cfg, options={'npoly': 15},
                                    priors={'teff': (5200., 300.)})
            out['c1_0_fix_vel'] = res['vel']
            out['c1_0_fix_param'] = np.array([res['param'][k] for k in synth.PARNAMES])
            out['c1_0_fix_chisq'] = res['chisq']
    fg = R.vel_fit.firstguess(specdata_of(R, objs[0]), config=cfg, options={'npoly': 15},
                              paramsgrid={'logg': [1, 3, 4.5], 'teff': [4000, 6000, 9000],
                                          'feh': [-1.5, -0.5], 'alpha': [0.2]},
                              vsinigrid=(None, 50))
    out['c1_fg'] = np.array([fg[k] for k in synth.PARNAMES] + [fg.get('vsini', -1)])
    np.savez_compressed(os.path.join(HERE, 'process.npz'), **out)


def gold_ccf(R):
    """fitter_ccf.fit on a Gaia-RVS-shaped window and on two arms, with the
    bank built by the reference's own make_ccf.preprocess_model_list."""
    out = {}
    for tag, shapes, every, npts in (('rvs', ('gaiarvs',), 9, 2048),
                                     ('two', ('desi_b', 'desi_r'), 12, 4096)):
        arms = [synth.make_setup(s, 'tiny', seed=31 + k) for k, s in enumerate(shapes)]
        cfg = frozen_config(R, max_vel=600 if tag == 'rvs' else 1000, vel_step0=2.5)
        CC = R.fitter_ccf.CCFCache
        vsinis = [0., 30., 300.] if tag == 'rvs' else [0., 100.]
        for a in arms:
            sh = synth.SHAPES[a['shape']]
            conf = R.make_ccf.get_ccf_config(logl0=np.log(sh['t_lo']),
                                             logl1=np.log(sh['t_hi']), npoints=npts)
            inds = np.arange(0, a['dats'].shape[0], every)
            specs = np.exp(a['dats'][inds].astype(np.float64))
            vec = a['vec'].T[inds].copy()
            vec[:, 0] = 10**vec[:, 0]
            models, params, vs = R.make_ccf.preprocess_model_list(a['lam'], specs, vec,
                                                                  conf, vsinis=vsinis)
            nm = a['name']
            CC.ccfs[nm] = np.fft.rfft(models, axis=1)
            CC.ccf2s[nm] = np.fft.rfft(models**2, axis=1)
            CC.ccf_models[nm] = models
            CC.ccf_info[nm] = dict(params=params, ccfconf=conf, vsinis=vs,
                                   parnames=a['parnames'])
            out[f'{tag}_{nm}_models'] = models.astype(np.float64)
            out[f'{tag}_{nm}_params'] = params
            out[f'{tag}_{nm}_vsinis'] = np.array(vs)
            out[f'{tag}_{nm}_conf'] = np.array([conf['logl0'], conf['logl1'],
                                                conf['npoints'], conf['splinestep']])
        objs = make_objects(arms, 'tiny', 2, 8100, sn_range=(30, 80), vel_sig=120.)
        pack_objects(objs, out, tag + '_')
        for i, o in enumerate(objs):
            sd = specdata_of(R, o)
            for a, s in enumerate(sd):
                ps, pi = R.make_ccf.preprocess_data(s.lam, s.spec, s.espec,
                                                    badmask=s.badmask,
                                                    ccfconf=CC.ccf_info[s.name]['ccfconf'])
                out[f'{tag}_{i}_{a}_proc_spec'] = ps
                out[f'{tag}_{i}_{a}_proc_ivar'] = pi
            res = R.fitter_ccf.fit(sd, cfg)
            out[f'{tag}_{i}_best_vel'] = res['best_vel']
            out[f'{tag}_{i}_best_vsini'] = res['best_vsini']
            out[f'{tag}_{i}_best_ccf'] = res['best_ccf']
            out[f'{tag}_{i}_best_par'] = np.array([res['best_par'][k] for k in synth.PARNAMES])
            out[f'{tag}_{i}_vel_grid'] = res['vel_grid']
    np.savez_compressed(os.path.join(HERE, 'ccf.npz'), **out)


def gold_switches(R):
    """The rarely used switches of get_chisq on the objects
    and evaluation points of chisq.npz: fast_interp (nearest-knot lookup instead of
    the spline, spec_fit.py:913-918), espec_systematic as a scalar and as a dictionary
    (spec_fit.py:933-940), outside_penalty=False (spec_fit.py:895-896)."""
    out = {}
    cfg = frozen_config(R)
    st = synth.make_setup('test', 'tiny', seed=3)
    inject_grid(R, st, 'test')
    objs = make_objects([st], 'tiny', 2, 500, bad_frac=0.02)
    prev = np.load(os.path.join(HERE, 'chisq.npz'))
    ev = prev['one_eval']
    opts = {'npoly': 15}
    res = {k: np.zeros((len(objs), len(ev))) for k in ('fast', 'sys_scalar', 'sys_dict', 'nopen')}
    for i, o in enumerate(objs):
        sd = specdata_of(R, o)
        assert np.array_equal(sd[0].spec, prev[f'one_{i}_0_spec'])
        sysv = 0.05 * float(np.median(sd[0].espec)) * 20
        out[f'sys_{i}'] = sysv
        for j, e in enumerate(ev):
            rot = None if e[5] < 0 else (e[5],)
            a = (sd, e[0], tuple(e[1:5]), rot)
            res['fast'][i, j] = R.spec_fit.get_chisq(*a, options=opts, config=cfg,
                                                     fast_interp=True)
            res['sys_scalar'][i, j] = R.spec_fit.get_chisq(*a, options=opts, config=cfg,
                                                           espec_systematic=sysv)
            res['sys_dict'][i, j] = R.spec_fit.get_chisq(*a, options=opts, config=cfg,
                                                         espec_systematic={'test': 2 * sysv})
            res['nopen'][i, j] = R.spec_fit.get_chisq(*a, options=opts, config=cfg,
                                                      outside_penalty=False)
    for k, v in res.items():
        out[k] = v
    np.savez_compressed(os.path.join(HERE, 'switches.npz'), **out)


def gold_resol(R):
    """Resolution-matrix mode (spec_fit.py:410-492,
    922-929): SpecData.resolution and the resol_params dictionary, on the one-arm
    objects / evaluation points of chisq.npz, on a three-arm DESI-shaped object with
    an 11-diagonal matrix per arm, through find_best and through a complete process()."""
    import scipy.sparse
    out = {}
    cfg = frozen_config(R)
    st = synth.make_setup('test', 'tiny', seed=3)
    inject_grid(R, st, 'test')
    objs = make_objects([st], 'tiny', 2, 500, bad_frac=0.02)
    prev = np.load(os.path.join(HERE, 'chisq.npz'))
    ev = prev['one_eval']
    opts = {'npoly': 15}
    out['one_R'] = np.array([1500., 4000.])
    chi = np.zeros((len(objs), len(ev)))
    chi_rp, chi_fast, chi_sys = np.zeros_like(chi), np.zeros_like(chi), np.zeros_like(chi)
    out['one_sys'] = np.zeros(len(objs))
    for i, o in enumerate(objs):
        nm, lam, sp, es, bad = o['arms'][0]
        assert np.array_equal(sp, prev[f'one_{i}_0_spec'])
        rm = R.spec_fit.construct_resol_mat(lam, resol=out['one_R'][i])
        dia = scipy.sparse.dia_matrix(rm.mat)
        out[f'one_{i}_offsets'] = dia.offsets
        out[f'one_{i}_data_sum'] = checksum(dia.data)
        if i == 0:
            out['one_0_data'] = dia.data
        sd_res = [R.spec_fit.SpecData(nm, lam, sp, es, badmask=bad, resolution=rm)]
        sd_plain = specdata_of(R, o)
        for j, e in enumerate(ev):
            rot = None if e[5] < 0 else (e[5],)
            chi[i, j] = R.spec_fit.get_chisq(sd_res, e[0], tuple(e[1:5]), rot, options=opts,
                                             config=cfg)
            chi_rp[i, j] = R.spec_fit.get_chisq(sd_plain, e[0], tuple(e[1:5]), rot,
                                                resol_params={'test': rm}, options=opts,
                                                config=cfg)
            # the matrix combined with the other switches of row a18
            out['one_sys'][i] = 0.05 * float(np.median(es)) * 20
            chi_fast[i, j] = R.spec_fit.get_chisq(sd_res, e[0], tuple(e[1:5]), rot, options=opts,
                                                  config=cfg, fast_interp=True)
            chi_sys[i, j] = R.spec_fit.get_chisq(sd_res, e[0], tuple(e[1:5]), rot, options=opts,
                                                 config=cfg, espec_systematic=out['one_sys'][i],
                                                 outside_penalty=False)
        e = ev[0]
        full = R.spec_fit.get_chisq(sd_res, e[0], tuple(e[1:5]), None if e[5] < 0 else (e[5],),
                                    options=opts, config=cfg, full_output=True)
        out[f'one_{i}_full_chisq'] = full['chisq']
        out[f'one_{i}_full_chisq_array'] = np.array(full['chisq_array'])
        out[f'one_{i}_full_model'] = full['models'][0]
        out[f'one_{i}_full_raw'] = full['raw_models'][0]
        vg = np.arange(-400, 400, 10.)
        plist = [tuple(o['params']), PROBE_PARAMS[1]]
        fb = R.spec_fit.find_best(sd_res, vg, plist, rot_params=(25.,), options=opts, config=cfg)
        for k in ('best_chi', 'best_vel', 'vel_err', 'kurtosis', 'skewness', 'probs'):
            out[f'one_{i}_fb_{k}'] = fb[k]
        out[f'one_{i}_fb_best_param'] = np.array(fb['best_param'])
    out['one_chisq'], out['one_chisq_resol_params'] = chi, chi_rp
    out['one_chisq_fast'], out['one_chisq_sys_nopen'] = chi_fast, chi_sys
    # three arms, fixed-width Gaussian matrices cut to 11 diagonals (the shape of DESI's
    # resolution data, desi_fit.py:723-748)
    arms = []
    for k, a in enumerate(('desi_b', 'desi_r', 'desi_z')):
        sa = synth.make_setup(a, 'tiny', seed=21 + k)
        inject_grid(R, sa, a)
        arms.append(sa)
    cfgd = frozen_config(R, min_vel=-1500, max_vel=1500)
    o = make_objects(arms, 'tiny', 1, 700, bad_frac=0.01)[0]
    assert np.array_equal(o['arms'][0][2], prev['desi_0_0_spec'])
    sds = []
    for a, (nm, lam, sp, es, bad) in enumerate(o['arms']):
        full = scipy.sparse.dia_matrix(R.spec_fit.construct_resol_mat(lam, width=0.9 + 0.2 * a).mat)
        keep = np.abs(full.offsets) <= 5
        m = scipy.sparse.dia_matrix((full.data[keep], full.offsets[keep]), shape=full.shape)
        out[f'desi_{a}_offsets'] = m.offsets
        out[f'desi_{a}_data'] = m.data.astype(np.float64)
        sds.append(R.spec_fit.SpecData(nm, lam, sp, es, badmask=bad,
                                       resolution=R.spec_fit.ResolMatrix(m)))
    dev = prev['desi_eval']
    out['desi_chisq'] = np.array([
        R.spec_fit.get_chisq(sds, e[0], tuple(e[1:5]), None if e[5] < 0 else (e[5],),
                             options={'npoly': 10}, config=cfgd) for e in dev])
    vg = np.arange(-1500, 1500, 5.)[::6]
    out['desi_scan_chisq'] = np.array([
        R.spec_fit.get_chisq(sds, v, tuple(o['params']), (12.,), options={'npoly': 10},
                             config=cfgd) for v in vg])
    # a complete fit with a resolution matrix (config-1 shape, object 0 of process.npz)
    st1 = synth.make_setup('test', 'test', seed=3)
    inject_grid(R, st1, 'test')
    o1 = make_objects([st1], 'test', 3, 4000, sn_range=(60, 150), vel_sig=100., bad_frac=0.0)[0]
    pp = np.load(os.path.join(HERE, 'process.npz'))
    nm, lam, sp, es, bad = o1['arms'][0]
    assert np.array_equal(sp, pp['c1_0_0_spec'])
    out['proc_R'] = 3000.
    rm = R.spec_fit.construct_resol_mat(lam, resol=3000.)
    sd1 = [R.spec_fit.SpecData(nm, lam, sp, es, badmask=bad, resolution=rm)]
    start = {'logg': 2, 'teff': 5000, 'feh': -0.2, 'alpha': 0.2, 'vsini': 0.1}
    res = R.vel_fit.process(sd1, dict(start), fixParam=[], config=cfg, options={'npoly': 15})
    for k in ('vel', 'vel_err', 'vel_skewness', 'vel_kurtosis', 'vsini', 'chisq', 'logl'):
        out[f'proc_{k}'] = res[k]
    out['proc_param'] = np.array([res['param'][k] for k in synth.PARNAMES])
    out['proc_param_err'] = np.array([res['param_err'][k] for k in synth.PARNAMES])
    out['proc_chisq_array'] = np.array(res['chisq_array'])
    out['proc_yfit'] = res['yfit'][0]
    out['proc_success'] = res['minimize_success']
    np.savez_compressed(os.path.join(HERE, 'resol.npz'), **out)


def gold_branches(R):
    """Branches the other fixtures do not reach:
      * a Gaia-RVS-shaped arm (0.05 A template sampling: 1.75 km/s per knot) evaluated at
        vsini up to 450 km/s -- rotation kernels of up to ~260 one-sided taps, beyond
        the fused path's limit of 128, and in between the case where only the ROUNDED
        tap bound exceeds it -- with find_best at vsini 300;
      * an object with RAGGED arms (one of three arms missing) at off-grid parameter
        vectors: the off-grid penalty is added once per arm the object has;
      * a complete vel_fit.process on a three-arm DESI-shaped object;
      * get_chisq on the FULL 28 600-node DESI layout (the bench's banks) at six points;
      * get_chisq_continuum for SpecData carrying a resolution matrix;
      * fitter_ccf.fit at npoints 8192 (make_ccf.preprocess_data outputs + first guess)."""
    import scipy.sparse
    out = {}
    cfg = frozen_config(R)
    # ---- Gaia-RVS, high vsini
    st = synth.make_setup('gaiarvs', 'tiny', seed=41)
    inject_grid(R, st, 'gaiarvs')
    out['gaia_dats_sum'] = checksum(st['dats'])
    objs = make_objects([st], 'tiny', 2, 9100, sn_range=(40, 120), vel_sig=60.)
    pack_objects(objs, out, 'gaia_')
    rs = np.random.RandomState(8)
    ev = []
    for k, vs in enumerate([-1, 5., 60., 150., 215., 300., 450., 300.]):
        p = synth.random_params('tiny', 1, 9200 + k)[0]
        ev.append([rs.uniform(-150, 150), *p, vs])
    ev = np.array(ev)
    out['gaia_eval'] = ev
    res = np.zeros((len(objs), len(ev)))
    for i, o in enumerate(objs):
        sd = specdata_of(R, o)
        for j, e in enumerate(ev):
            rot = None if e[5] < 0 else (e[5],)
            res[i, j] = R.spec_fit.get_chisq(sd, e[0], tuple(e[1:5]), rot,
                                             options={'npoly': 10}, config=cfg)
    out['gaia_chisq'] = res
    vg = np.arange(-500, 500, 5.)
    sd = specdata_of(R, objs[0])
    fb = R.spec_fit.find_best(sd, vg, [tuple(objs[0]['params'])], rot_params=(300.,),
                              options={'npoly': 10}, config=cfg)
    out['gaia_scan_grid'] = vg
    for k in ('best_chi', 'best_vel', 'vel_err', 'kurtosis', 'skewness'):
        out[f'gaia_scan_{k}'] = fb[k]
    # ---- CCF at 8192 points: preprocessing outputs and the first guess (the bank is
    # rebuilt by the test from the same seeds with this package's builder; its
    # continuum-normalised models are pinned by ccf.npz at smaller sizes)
    sh = synth.SHAPES['gaiarvs']
    conf = R.make_ccf.get_ccf_config(logl0=np.log(sh['t_lo']), logl1=np.log(sh['t_hi']),
                                     npoints=8192)
    out['gaia_ccf_conf'] = np.array([conf['logl0'], conf['logl1'], conf['npoints'],
                                     conf['splinestep']])
    inds = np.arange(0, st['dats'].shape[0], 3)
    specs = np.exp(st['dats'][inds].astype(np.float64))
    vec = st['vec'].T[inds].copy()
    vec[:, 0] = 10**vec[:, 0]
    models, params, vs = R.make_ccf.preprocess_model_list(st['lam'], specs, vec, conf,
                                                          vsinis=[0., 100., 300.])
    CC = R.fitter_ccf.CCFCache
    CC.ccfs['gaiarvs'] = np.fft.rfft(models, axis=1)
    CC.ccf2s['gaiarvs'] = np.fft.rfft(models**2, axis=1)
    CC.ccf_models['gaiarvs'] = models
    CC.ccf_info['gaiarvs'] = dict(params=params, ccfconf=conf, vsinis=vs,
                                  parnames=st['parnames'])
    out['gaia_ccf_params'], out['gaia_ccf_vsinis'] = params, np.array(vs)
    out['gaia_ccf_models_sum'] = models.sum(axis=1)
    ccfg = frozen_config(R, max_vel=600, vel_step0=2.5)
    for i, o in enumerate(objs):
        sd = specdata_of(R, o)
        ps, pi = R.make_ccf.preprocess_data(sd[0].lam, sd[0].spec, sd[0].espec,
                                            badmask=sd[0].badmask, ccfconf=conf)
        out[f'gaia_ccf_{i}_proc_spec'], out[f'gaia_ccf_{i}_proc_ivar'] = ps, pi
        r = R.fitter_ccf.fit(sd, ccfg)
        out[f'gaia_ccf_{i}_best_vel'] = r['best_vel']
        out[f'gaia_ccf_{i}_best_vsini'] = r['best_vsini']
        out[f'gaia_ccf_{i}_best_par'] = np.array([r['best_par'][k] for k in synth.PARNAMES])
        out[f'gaia_ccf_{i}_best_ccf_min'] = float(np.min(r['best_ccf']))
    # ---- three DESI arms (tiny layout): ragged arms off the grid, and a complete fit
    names = ('desi_b', 'desi_r', 'desi_z')
    arms = [synth.make_setup(s, 'tiny', seed=21 + k) for k, s in enumerate(names)]
    for a in arms:
        inject_grid(R, a)
    out['desi_dats_sum'] = np.array([checksum(a['dats']) for a in arms])
    objs3 = make_objects(arms, 'tiny', 2, 7300, sn_range=(30, 90), vel_sig=120.)
    pack_objects(objs3, out, 'd3_')
    offgrid = np.array([[35., 9500., 2.0, -1.0, 0.2, 8.],      # teff above the grid
                        [-20., 5000., 2.0, -2.4, 0.2, -1],      # feh below
                        [5., 5000., 5.4, -1.0, 1.3, 20.],       # logg and alpha above
                        [60., 4800., 2.2, -0.8, 0.4, 12.]])     # inside
    out['d3_offgrid_eval'] = offgrid
    rag = np.zeros((3, len(offgrid)))
    sd_all = specdata_of(R, objs3[0])
    for r, keep in enumerate(((0, 1, 2), (0, 2), (1,))):
        sd = [sd_all[k] for k in keep]
        for j, e in enumerate(offgrid):
            rot = None if e[5] < 0 else (e[5],)
            rag[r, j] = R.spec_fit.get_chisq(sd, e[0], tuple(e[1:5]), rot,
                                             options={'npoly': 10}, config=cfg)
    out['d3_ragged_chisq'] = rag
    start = {'teff': 5500., 'logg': 3.0, 'feh': -1.0, 'alpha': 0.3, 'vsini': 10.}
    keys = ('vel', 'vel_err', 'vel_skewness', 'vel_kurtosis', 'vsini', 'chisq')
    for i, o in enumerate(objs3):
        res3 = R.vel_fit.process(specdata_of(R, o), dict(start), fixParam=[], config=cfg,
                                 options={'npoly': 10})
        for k in keys:
            out[f'd3_{i}_{k}'] = res3[k]
        out[f'd3_{i}_param'] = np.array([res3['param'][k] for k in synth.PARNAMES])
        out[f'd3_{i}_param_err'] = np.array([res3['param_err'][k] for k in synth.PARNAMES])
        out[f'd3_{i}_chisq_array'] = np.array(res3['chisq_array'])
        out[f'd3_{i}_success'] = res3['minimize_success']
    # ---- continuum-only fit with a resolution matrix (spec_fit.py:739-783)
    sd0 = sd_all[0]
    rm = R.spec_fit.construct_resol_mat(sd0.lam, width=1.1)
    sdr = R.spec_fit.SpecData(sd0.name, sd0.lam, sd0.spec, sd0.espec, badmask=sd0.badmask,
                              resolution=rm)
    cc = R.spec_fit.get_chisq_continuum([sdr, sd_all[1]], options={'npoly': 10})
    out['cont_resol_chisq'] = np.array(cc['chisq_array'])
    out['cont_resol_redchisq'] = np.array(cc['redchisq_array'])
    # ---- the full DESI layout (28 600 nodes): the banks of bench.py
    sys.path.insert(0, ROOT)
    import bench
    setups, objects, pars, vel = bench.make_inputs('desi', 4, 4242)
    out['full_dats_sum'] = np.array([checksum(s['dats'][::97]) for s in setups])
    for s in setups:
        inject_grid(R, s)
    tp, tv, tvs = bench.trial_points(pars, vel, 'desi', 3, 77)
    cfgd = R.utils.freezeDict(bench.make_config(bench.WORKLOADS['desi']))
    full = np.zeros((2, 3))
    for i in range(2):
        sd = [R.spec_fit.SpecData(a[0], a[1], a[2], a[3], badmask=a[4]) for a in objects[i]]
        for e in range(3):
            full[i, e] = R.spec_fit.get_chisq(sd, tv[e, i], tuple(tp[e, i]), (tvs[e, i],),
                                              options={'npoly': 10}, config=cfgd)
    out['full_chisq'] = full
    np.savez_compressed(os.path.join(HERE, 'branches.npz'), **out)


def gold_specdata(R):
    """desi_fit.get_specdata (desi/desi_fit.py:781-888): masking, bridging, error clamp,
    dichroic mask, and the deconvolved resolution matrices, on synthetic DESI-like frames
    with zero / negative / non-finite inverse variances, masked runs at the ends and in
    the middle, an all-masked arm and a zero-median arm."""
    import importlib
    for m in ('astropy.table', 'astropy.units'):
        sys.modules.setdefault(m, types.ModuleType(m))
    desi_fit = importlib.import_module('rvspecfit.desi.desi_fit')
    out = {}
    rs = np.random.RandomState(12)
    setups = ['b', 'r', 'z']
    waves = {'b': np.arange(4000., 4700.1, 0.8), 'r': np.arange(6000., 6400.1, 0.8),
             'z': np.arange(8000., 8400.1, 0.8)}
    nfib, width = 5, 11
    fluxes, ivars, masks, resol = {}, {}, {}, {}
    for s in setups:
        n = len(waves[s])
        fl = 20 * (1 + 0.3 * np.sin(waves[s] / 200.))[None, :] + rs.normal(size=(nfib, n))
        iv = 1. / (0.5 + rs.uniform(size=(nfib, n)))**2
        mk = np.zeros((nfib, n), dtype=int)
        iv[0, 100:130] = 0
        iv[0, 400] = -1
        iv[1, :12] = 0
        iv[1, -7:] = 0
        fl[1, 250] = np.nan
        iv[2, 300] = np.inf
        mk[2, 330:360] = 4
        iv[3, rs.choice(n, 30, replace=False)] = 1e6         # errors to be clamped
        xs = np.arange(width) - width // 2
        sig = 1.2 + 0.3 * np.sin(np.arange(n) / 300.)
        band = np.exp(-0.5 * (xs[:, None] / sig[None, :])**2)
        resol[s] = np.tile((band / band.sum(axis=0))[None], (nfib, 1, 1))
        out[f'resol_{s}'] = resol[s][0]                     # the same band for every fibre
        fluxes[s], ivars[s], masks[s] = fl, iv, mk
    masks['r'][4, :] = 1                                    # arm dropped
    fluxes['z'][4] = np.where(rs.uniform(size=len(waves['z'])) < 0.6, 0.0, fluxes['z'][4])
    for s in setups:
        out[f'wave_{s}'], out[f'flux_{s}'], out[f'ivar_{s}'] = waves[s], fluxes[s], ivars[s]
        out[f'mask_{s}'] = masks[s]
    sig0 = {'b': 0.5, 'r': 0.45, 'z': 0.4}
    for mode, kw in (('plain', {}), ('resol', dict(use_resolution_matrix=True,
                                                      lsf_sigma0_angstrom=sig0))):
        for f in range(nfib):
            sds = desi_fit.get_specdata(waves, fluxes, ivars, masks, resol, f, setups, **kw)
            out[f'{mode}_{f}_names'] = np.array([sd.name for sd in sds] if sds else [])
            for sd in sds or ():
                out[f'{mode}_{f}_{sd.name}_spec'] = sd.spec
                out[f'{mode}_{f}_{sd.name}_espec'] = sd.espec
                out[f'{mode}_{f}_{sd.name}_bad'] = sd.badmask
                if sd.resolution is not None and f in (0, 3):
                    dia = scipy.sparse.dia_matrix(sd.resolution.mat)
                    out[f'{mode}_{f}_{sd.name}_resol_offsets'] = dia.offsets
                    out[f'{mode}_{f}_{sd.name}_resol_data'] = dia.data
    np.savez_compressed(os.path.join(HERE, 'specdata.npz'), **out)


ALL = dict(specdata=gold_specdata, kat=gold_kat, interp=gold_interp, chisq=gold_chisq, process=gold_process,
           ccf=gold_ccf, switches=gold_switches, resol=gold_resol, branches=gold_branches)


if __name__ == '__main__':
    which = sys.argv[1:] or list(ALL)
    R = load_reference()
    for w in which:
        print('generating', w, flush=True)
        ALL[w](R)
    print('done')
