"""Host logic of the batched drivers, no GPU: the lock-step Nelder-Mead and the
lock-step BFGS must follow scipy's own trajectories bit for bit on objectives
that are cheap to evaluate on the CPU, and the pipelined process_batch must
reproduce vel_fit.process object by object."""
import numpy as np
import scipy.optimize

from rvspecfit_b200 import batch_fit


def _problems(B, N, seed):
    rs = np.random.RandomState(seed)
    cen = rs.normal(size=(B, N))
    scale = np.exp(rs.normal(size=(B, N)))
    rot = rs.normal(size=(B, N, N)) * 0.3

    def f_one(b, x):
        d = (x - cen[b]) * scale[b]
        d = d + rot[b] @ d
        return float(np.sum(d**2) + 0.05 * np.sum(np.abs(d)**1.5) + 0.3 * np.sin(3 * d[0]))

    def fbatch(idx, X):
        return np.array([f_one(int(b), x) for b, x in zip(idx, X)])
    return f_one, fbatch


def test_lockstep_nelder_mead_is_scipy():
    B, N = 24, 5
    f_one, fbatch = _problems(B, N, 3)
    rs = np.random.RandomState(7)
    sims = rs.normal(size=(B, N + 1, N)) * 2
    res = batch_fit.nelder_mead_lockstep(fbatch, sims, xatol=1e-2, fatol=1e-3, maxiter=10000)
    few = batch_fit.nelder_mead_lockstep(fbatch, sims, xatol=1e-9, fatol=1e-12, maxiter=40)
    spec = batch_fit.nelder_mead_lockstep(fbatch, sims, xatol=1e-2, fatol=1e-3, maxiter=10000,
                                          speculate_below=10)
    for k in ('x', 'fun', 'final_simplex', 'nit', 'nfev', 'success'):
        assert np.array_equal(spec[k], res[k]), k
    # independent lock-step sets with deferred evaluation: same trajectories
    order = []

    def fsubmit(idx, X):
        order.append(len(idx))
        return lambda: fbatch(idx, X)
    for groups in (1, 2, 5):
        inter = batch_fit.nelder_mead_interleaved(fsubmit, sims, groups=groups, xatol=1e-2,
                                                  fatol=1e-3, maxiter=10000, speculate_below=4)
        for k in ('x', 'fun', 'final_simplex', 'nit', 'nfev', 'success'):
            assert np.array_equal(inter[k], res[k]), (groups, k)
    # the library's host-side stepper (csrc/nm_host.cpp): the same trajectories
    for kw, ref in ((dict(), res), (dict(speculate_below=10), res), (dict(speculate_below=10**6), res),
                    (dict(xatol=1e-9, fatol=1e-12, maxiter=40), few)):
        nat = batch_fit.nelder_mead_lockstep(fbatch, sims, native=True,
                                             **{**dict(xatol=1e-2, fatol=1e-3, maxiter=10000), **kw})
        for k in ('x', 'fun', 'final_simplex', 'nit', 'nfev', 'success'):
            assert np.array_equal(nat[k], ref[k]), (kw, k)
    for b in range(B):
        opts = {'fatol': 1e-3, 'xatol': 1e-2, 'initial_simplex': sims[b], 'maxiter': 10000,
                'maxfev': np.inf}
        want = scipy.optimize.minimize(lambda x: f_one(b, x), sims[b, 0], method='Nelder-Mead',
                                       options=opts)
        assert np.array_equal(res['x'][b], want['x']), b
        assert res['fun'][b] == want['fun'] and res['success'][b] == want['success']
        assert np.array_equal(res['final_simplex'][b], want['final_simplex'][0])
        assert res['nit'][b] == want['nit'] and res['nfev'][b] == want['nfev']
        opts.update(fatol=1e-12, xatol=1e-9, maxiter=40)
        want = scipy.optimize.minimize(lambda x: f_one(b, x), sims[b, 0], method='Nelder-Mead',
                                       options=opts)
        assert not want['success'] and not few['success'][b]
        assert np.array_equal(few['final_simplex'][b], want['final_simplex'][0])


def _noisy_problems(B, N, seed, noise):
    """Smooth problems with values ~1e4 plus a rapidly varying term of the given
    amplitude: the forward-difference gradient (step 1.5e-8) sees it as noise, which
    drives scipy's BFGS through the fallback line search and the precision-loss exit
    the way chi-square surfaces do."""
    f_smooth, _ = _problems(B, N, seed)

    def f_one(b, x):
        return 1e4 + f_smooth(b, x) + noise * np.sin(1e9 * x[0] + 3e8 * x[-1])

    def fbatch(idx, X):
        return np.array([f_one(int(b), x) for b, x in zip(idx, X)])
    return f_one, fbatch


def test_lockstep_bfgs_is_scipy():
    import warnings
    from rvspecfit_b200 import batch_bfgs
    seen = set()
    for B, N, seed, noise in ((24, 4, 11, 0.0), (24, 6, 5, 1e-7), (16, 6, 6, 1e-9),
                              (16, 3, 7, 1e-5), (8, 6, 8, 1e-3), (8, 2, 9, 0.0)):
        f_one, fbatch = _noisy_problems(B, N, seed, noise)
        rs = np.random.RandomState(seed + 1)
        x0 = rs.normal(size=(B, N))
        H0 = np.diag(np.exp(rs.normal(size=N) * 2))
        calls = []

        def counted(idx, X):
            calls.append(len(idx))
            assert len(idx) % (N + 1) == 0
            return fbatch(idx, X)
        got = batch_bfgs.bfgs_lockstep(counted, x0, H0)
        for b in range(B):
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                want = scipy.optimize.minimize(lambda x: f_one(b, x), x0[b], method='BFGS',
                                               options=dict(hess_inv0=H0))
            assert np.array_equal(got['x'][b], want['x']), (noise, b)
            assert got['fun'][b] == want['fun'] and got['nit'][b] == want['nit']
            assert got['status'][b] == want['status']
            seen.add(int(want['status']))
        # requests were gathered: far fewer batched calls than scalar evaluations
        assert len(calls) == got['rounds'] and len(calls) < sum(calls) / 4
    assert seen == {0, 2}      # both the converged and the precision-loss exit were met


def test_native_bfgs_follows_the_lockstep_restatement():
    """BFGSStepper (csrc/bfgs_host.cpp) against batch_bfgs.bfgs_steps, which is scipy bit for
    bit: same algorithm, matrix products summed in index order instead of through BLAS.
    The forward-difference gradient of a function of size 1e4 carries rounding noise of
    ~1e-4 that depends on the last bits of x, so trajectories separate after a few
    iterations whatever the implementation; the check is therefore: identical first
    iteration, identical decisions (iteration counts, exits, rounds -- through the MINPACK
    search, the fallback search and its zoom phase) while the paths are still close, and
    the same exits, values and iteration counts in distribution at the end."""
    from rvspecfit_b200 import batch_bfgs
    for B, N, seed, noise in ((24, 4, 11, 0.0), (8, 2, 9, 0.0), (24, 6, 5, 1e-7), (16, 6, 6, 1e-9),
                              (16, 3, 7, 1e-5), (8, 6, 8, 1e-3)):
        f_one, fbatch = _noisy_problems(B, N, seed, noise)
        rs = np.random.RandomState(seed + 1)
        x0 = rs.normal(size=(B, N))
        H0 = np.diag(np.exp(rs.normal(size=N) * 2))
        for maxiter in (1, 2):
            want = batch_bfgs.bfgs_lockstep(fbatch, x0, H0, maxiter=maxiter)
            got = batch_bfgs.bfgs_lockstep(fbatch, x0, H0, native=True, maxiter=maxiter)
            assert got['rounds'] == want['rounds'], (noise, maxiter)
            assert np.array_equal(got['nit'], want['nit'])
            assert np.array_equal(got['status'], want['status'])
            if maxiter == 1:
                assert np.allclose(got['x'], want['x'], rtol=0, atol=1e-13)
                assert np.allclose(got['fun'], want['fun'], rtol=1e-15)
        want = batch_bfgs.bfgs_lockstep(fbatch, x0, H0)
        calls = []

        def counted(idx, X):
            calls.append(len(idx))
            assert len(idx) % (N + 1) == 0
            return fbatch(idx, X)
        got = batch_bfgs.bfgs_lockstep(counted, x0, H0, native=True)
        assert len(calls) == got['rounds']
        assert np.mean(got['status'] == want['status']) >= 0.75
        if noise == 0.0:
            assert np.all(np.abs(got['fun'] - want['fun']) <= 1e-9 * np.abs(want['fun']))
        # the same descent on average (a noisy gradient stalls both wherever it happens to)
        f0 = fbatch(np.arange(B), x0)
        dec_got, dec_want = np.mean(f0 - got['fun']), np.mean(f0 - want['fun'])
        assert 0.8 <= dec_got / dec_want <= 1.25, (noise, dec_got, dec_want)
        assert abs(np.mean(got['nit']) - np.mean(want['nit'])) <= 0.3 * np.mean(want['nit']) + 2


def test_native_bfgs_edge_cases_follow_the_restatement():
    """Starts that are already stationary, objectives that turn NaN or infinite, a single
    iteration budget: the exits (scipy's warnflag) and iteration counts of BFGSStepper equal
    those of bfgs_steps; and while some problems still search, result() already holds the
    final rows of those that have stopped (what the hand-over between stages relies on)."""
    from rvspecfit_b200 import batch_bfgs
    B, N = 12, 3
    cen = np.linspace(-1, 1, B * N).reshape(B, N)

    def fbatch(idx, X):
        d = X - cen[idx]
        f = 5.0 + np.sum(d**2, axis=1) + 0.1 * np.sum(d**4, axis=1)
        f = np.where(X[:, 0] > 2.5, np.nan, f)            # a wall of NaN
        f = np.where(X[:, 1] < -2.5, np.inf, f)           # and one of +inf
        return f
    x0 = cen + 0.3
    x0[0] = cen[0]                      # stationary start: no iteration
    x0[1] = cen[1] + [3.0, 0, 0]        # starts behind the NaN wall
    x0[2] = cen[2] + [0, -3.2, 0]       # starts behind the inf wall
    x0[3] = cen[3] + [2.4 - cen[3, 0], 0, 0]    # next to the NaN wall, gradient pointing away
    for maxiter in (None, 1, 3):
        want = batch_bfgs.bfgs_lockstep(fbatch, x0, None, maxiter=maxiter)
        got = batch_bfgs.bfgs_lockstep(fbatch, x0, None, native=True, maxiter=maxiter)
        assert np.array_equal(got['status'], want['status']), (maxiter, got['status'], want['status'])
        assert np.array_equal(got['nit'], want['nit']), maxiter
        assert got['rounds'] == want['rounds']
        ok = np.isfinite(want['fun'])
        assert np.array_equal(np.isfinite(got['fun']), ok)
        assert np.allclose(got['x'][ok], want['x'][ok], rtol=0, atol=1e-7)
        assert np.allclose(got['fun'][ok], want['fun'][ok], rtol=1e-12)
    assert {0, 3} <= set(want['status'].tolist()) or {0, 2} <= set(want['status'].tolist())
    # rows of stopped problems are final while the others search
    st = batch_bfgs.BFGSStepper(x0 * np.linspace(1, 3, B)[:, None])
    snaps = []
    try:
        while True:
            req = st.request()
            if req is None:
                break
            st.feed(fbatch(*req))
            act = st.active()
            if act.any() and not act.all():
                snaps.append((~act, st.result()))
        final = st.result()
    finally:
        st.close()
    assert len(snaps) > 2
    for stopped, res in snaps:
        for k in ('x', 'fun', 'nit', 'status'):
            assert np.array_equal(res[k][stopped], final[k][stopped], equal_nan=True), k


def test_hessian_points_replay_central_hessian():
    """The vectorised Hessian stencil asks for exactly the points of
    vel_fit.central_hessian and combines their values with its arithmetic."""
    from rvspecfit_b200 import vel_fit
    rs = np.random.RandomState(2)
    B, n = 5, 4
    x = rs.normal(size=(B, n)) * [300, 1, 1, 0.3] + [5000, 3, -1, 0.2]
    hs = [vel_fit.HESS_STEP[k] for k in ('teff', 'logg', 'feh', 'alpha')]
    A = rs.normal(size=(B, n, n))

    def f(b, p):
        d = (p - x[b]) / [100, 0.3, 0.2, 0.1] + 0.3
        return float(d @ (A[b] @ A[b].T) @ d + np.sum(np.cos(d)))
    P = batch_fit.hessian_points(x, hs)
    vals = np.array([[f(b, p) for p in P[b]] for b in range(B)])
    H = batch_fit.hessian_from_values(vals, hs)
    for b in range(B):
        asked = []

        def rec(p):
            asked.append(np.array(p))
            return f(b, p)
        want = vel_fit.central_hessian(rec, x[b], hs)
        assert np.array_equal(H[b], want), b
        uniq = {p.tobytes() for p in asked}
        assert uniq == {p.tobytes() for p in P[b]}


class _AnalyticEngine:
    """Stand-in for LikelihoodEngine on the CPU: a smooth -2 log L per object."""
    NSLOT = 8

    def __init__(self, nobj, seed):
        rs = np.random.RandomState(seed)
        self.nobj = nobj
        self.objects = [[_Arm()] for _ in range(nobj)]
        self.truth = np.column_stack([rs.normal(0, 100, nobj), rs.uniform(3, 30, nobj),
                                      rs.uniform(4500, 6500, nobj), rs.uniform(1, 4.5, nobj),
                                      rs.uniform(-2, 0, nobj), rs.uniform(0, 0.4, nobj)])
        self.scale = np.array([3.0, 4.0, 80.0, 0.2, 0.1, 0.08])
        self.timer = None
        self.calls = []

    def chi(self, obj, vel, params, vsini):
        vel = np.asarray(vel, dtype=np.float64)
        z = [(vel - self.truth[obj, 0].reshape((-1,) + (1,) * (vel.ndim - 1))) / self.scale[0]]
        v = np.zeros(len(obj)) if vsini is None else np.asarray(vsini)
        z.append(((v - self.truth[obj, 1]) / self.scale[1]).reshape((-1,) + (1,) * (vel.ndim - 1)))
        for j in range(4):
            z.append(((params[:, j] - self.truth[obj, 2 + j]) / self.scale[2 + j]).reshape(
                (-1,) + (1,) * (vel.ndim - 1)))
        tot = 7000.0 + 40 * obj.reshape((-1,) + (1,) * (vel.ndim - 1))
        tot = tot + 300 * (1 - 1 / (1 + z[0]**2 / 300)) * 300
        for a, zz in enumerate(z[1:]):
            tot = tot + zz**2 + 0.1 * np.cos(zz + a) + 0.05 * zz * z[(a + 2) % 6]
        return tot

    def evaluate(self, obj, vels, params, vsini=None, want_model=False, **kw):
        obj = np.asarray(obj, dtype=np.int64)
        params = np.array(params, dtype=np.float64, ndmin=2)
        self.calls.append(len(obj))
        out = self.chi(obj, vels, params, vsini)
        if want_model:
            n = len(obj)
            ex = dict(moff=np.arange(n + 1) * 3, model=np.ones(3 * n), raw=np.ones(3 * n))
            return out, dict(arms={'fake': dict(sel=np.arange(n), extras=ex,
                                                tbad=np.zeros(n, dtype=bool))})
        return out

    def drain(self):
        pass


class _Arm:
    name = 'fake'
    lam = np.arange(3.0)
    spec = np.ones(3)
    espec = np.ones(3)
    badmask = np.zeros(3, dtype=bool)


def _host_scan_stats(vel_grid, chisq, quadratic=True, nv=None, want_probs=True):
    """spec_fit.scan_stats on the host (the oracle's restatement of find_best's tail)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), 'oracle'))
    import oracle
    out = np.zeros((len(vel_grid), 8))
    for s in range(len(vel_grid)):
        n = len(vel_grid[s]) if nv is None else int(nv[s])
        r = oracle.scan_statistics(vel_grid[s][:n], chisq[s][:, :n].T, quadratic)
        out[s, :5] = r['best_chi'], r['best_vel'], r['vel_err'], r['skewness'], r['kurtosis']
    return out, None


def test_process_batch_pipeline_equals_single_object_rules(monkeypatch):
    """process_batch (grouped coroutines, lock-step NM and BFGS, vectorised stencils)
    against vel_fit.process driven by scipy, on an analytic likelihood: every number
    of every object identical, whatever the grouping."""
    from rvspecfit_b200 import spec_fit, spec_inter, vel_fit
    B = 7
    eng = _AnalyticEngine(B, 4)
    names = ('teff', 'logg', 'feh', 'alpha')
    cfg = dict(min_vel=-1000, max_vel=1000, vel_step0=5, max_vsini=500, min_vsini=0.1,
               min_vel_step=0.2, second_minimizer=True, template_lib='none/')
    monkeypatch.setattr(spec_inter, 'getSpecParams', lambda setup, config: names)
    monkeypatch.setattr(spec_fit, 'scan_stats', _host_scan_stats)
    starts = [dict(teff=5500., logg=3., feh=-1., alpha=0.2, vsini=10.) for _ in range(B)]
    priors = {'teff': (5600., 400.)}
    res = {g: batch_fit.process_batch(None, starts, config=cfg, options={}, engine=eng,
                                      priors=priors, groups=g) for g in (1, 3)}
    # finished objects handed on in groups while the slower ones iterate (spawned
    # coroutines), by the calling thread and by one thread per coroutine
    monkeypatch.setattr(batch_fit, 'PEEL_MIN', 2)
    monkeypatch.setattr(batch_fit, 'MAX_CALL_ITEMS', 100)
    for key, kw in (('peel', dict(groups=1, threads=False)), ('peel-threads', dict(groups=2, threads=True))):
        n0 = len(eng.calls)
        res[key] = batch_fit.process_batch(None, starts, config=cfg, options={}, engine=eng,
                                           priors=priors, peel=True, **kw)
        assert all(r is not None for r in res[key])
        assert max(eng.calls[n0:]) <= 600     # scans of 7 objects at most; stencils in pieces
    assert batch_fit.process_batch.last_spawned > 0

    # the single-object path with the same likelihood behind the reference-shaped calls
    def get_chisq(specdata, vel, atm, rot=None, resol=None, options=None, config=None,
                  full_output=False, outside_penalty=True):
        i = np.array([specdata[0].index])
        c = float(eng.chi(i, np.array([float(vel)]), np.array([atm], dtype=np.float64),
                          None if rot is None else np.array([rot[0]]))[0])
        if full_output:
            return dict(chisq=c, logl=-0.5 * c, chisq_array=[0.0], npix_array=[3],
                        models=[np.ones(3)], raw_models=[np.ones(3)])
        return c

    def find_best(specdata, vel_grid, params_list, rot_params=None, resol_params=None,
                  options=None, config=None, quadratic=True):
        i = np.array([specdata[0].index])
        chi = eng.chi(i, np.asarray(vel_grid)[None, :], np.array(params_list, dtype=np.float64),
                      None if rot_params is None else np.array([rot_params[0]]))
        o = _host_scan_stats(np.asarray(vel_grid)[None], chi[None])[0][0]
        return dict(best_chi=o[0], best_vel=o[1], vel_err=o[2], skewness=o[3], kurtosis=o[4],
                    best_param=params_list[0])
    monkeypatch.setattr(spec_fit, 'get_chisq', get_chisq)
    monkeypatch.setattr(spec_fit, 'find_best', find_best)
    monkeypatch.setattr(spec_fit, 'param_dict_to_tuple',
                        lambda d, setup, config: tuple(d[k] for k in names))
    for i in range(B):
        sd = spec_fit.SpecData('fake', np.arange(3.0) + 1, np.ones(3), np.ones(3))
        sd.index = i
        want = vel_fit.process([sd], dict(starts[i]), config=cfg, options={}, priors=priors)
        for g in res:
            got = res[g][i]
            for k in ('vel', 'vel_err', 'vel_skewness', 'vel_kurtosis', 'vsini', 'chisq',
                      'minimize_success', 'bad_hessian'):
                assert got[k] == want[k], (g, i, k, got[k], want[k])
            for k in names:
                assert got['param'][k] == want['param'][k], (g, i, k)
                assert got['param_err'][k] == want['param_err'][k] or \
                    (np.isnan(got['param_err'][k]) and np.isnan(want['param_err'][k])), (g, i, k)
            assert np.array_equal(got['param_covar'], want['param_covar'])


def test_batch_objective_matches_scalar_rules():
    """Wall, vsini penalty and parameter layout of vel_fit.chisq_func, vectorised."""
    from rvspecfit_b200 import vel_fit

    class FakeEngine:
        def evaluate(self, idx, vel, params, vsini):
            v = 0 if vsini is None else vsini
            return vel * 1e-3 + params.sum(axis=1) + 10 * v + idx
    cfg = dict(min_vel=-1000, max_vel=1000, max_vsini=500)
    names = ('teff', 'logg', 'feh', 'alpha')
    p0 = [dict(teff=5000. + i, logg=2., feh=-1., alpha=0.2, vsini=3.) for i in range(3)]
    fobj = batch_fit.BatchObjective(FakeEngine(), names, p0, ['alpha'], True, cfg,
                                    priors={'teff': (5100., 200.)})
    X = np.array([[10., 5., 5200., 2.5, -0.5], [2000., 5., 5200., 2.5, -0.5],
                  [10., -2., 5200., 2.5, -0.5], [10., 600., 5200., np.nan, -0.5]])
    idx = np.array([0, 1, 2, 1])
    got = fobj(idx, X)
    vm = vel_fit.VSiniMapper(500)
    for k in range(4):
        pm = vel_fit.ParamMapper(names, p0[idx[k]], ['alpha'], vm, fitVsini=True)
        pd = pm.forward(X[k])
        if pd['vel'] > 1000 or pd['vel'] < -1000 or (~np.isfinite(pd['params'])).any():
            want = 1e30
        else:
            par = np.array(pd['params'])
            want = ((5100. - par[0]) / 200.)**2 + \
                (pd['vel'] * 1e-3 + par.sum() + 10 * pd['vsini'] + idx[k]) + pd['penalty']
        assert got[k] == want, k


def test_fit_pack_collect_equal_numpy_rules():
    """rvs_fit_pack / rvs_fit_collect (csrc/fit_host.cpp) against the numpy restatement of
    the same rules (BatchObjective.unpack / prior_term, LikelihoodEngine._collect_fast):
    identical bits for every fitted-vector layout."""
    import ctypes
    from rvspecfit_b200 import _cabi, _dev, spec_inter
    L = _cabi.lib()
    rs = np.random.RandomState(0)
    nobj, ns, narm = 9, 4, 3

    class Bank:
        log_ids = [0]
        ndim = 4

    class Eng:
        setups = ['a', 'b', 'c']
        arms = {'a': {'bank': Bank()}}

        def submit_fit(self, *a):
            return None
    eng = Eng()
    eng.nobj = nobj
    eng._oix = rs.randint(-1, 5, size=(narm, nobj)).astype(np.int32)
    eng.badchi = rs.uniform(100, 1000, nobj)
    eng._cover0 = rs.rand(nobj) > 0.2
    names = ['teff', 'logg', 'feh', 'alpha']
    p0 = [dict(teff=5000. + 100 * i, logg=2. + .1 * i, feh=-1., alpha=0.2, vsini=3. + i)
          for i in range(nobj)]
    cfg = dict(min_vel=-1000, max_vel=1000, max_vsini=500)
    hp = _dev.hptr
    for fix, fitv, pri in ((['alpha'], True, {'teff': (5100., 200.), 'feh': (-1., .3)}),
                           ([], True, None), (['teff'], False, {'logg': (2., 1.)}),
                           ([], False, None)):
        fobj = batch_fit.BatchObjective(eng, names, [dict(d) for d in p0],
                                        fix + ([] if fitv else ['vsini']), fitv, cfg, pri)
        lay = fobj.layout()
        assert lay
        K, Kp = 50, 64
        idx = rs.randint(0, nobj, K).astype(np.int32)
        cols = [rs.normal(0, 600, K)] + ([rs.normal(100, 300, K)] if fitv else []) + \
            [rs.uniform(4000, 7000, K) if n == 'teff' else rs.normal(0, 1, K)
             for n in names if n not in fix]
        X = np.ascontiguousarray(np.column_stack(cols))
        assert X.shape[1] == lay.nfit
        X[3, -1] = np.nan
        logvals = np.ascontiguousarray(np.log10(X[:, fobj._logcols].T)) if fobj._logcols else None
        h_in = np.zeros((2 + ns, Kp))
        h_oix = np.zeros((narm, Kp), dtype=np.int32)
        prior, pen, wall = np.zeros(K), np.zeros(K), np.zeros(K, dtype=np.uint8)
        vm = ctypes.c_double()
        rc = L.rvs_fit_pack(ctypes.byref(lay), K, Kp, hp(idx), hp(X),
                            None if logvals is None else hp(logvals), hp(h_in), hp(h_oix),
                            hp(prior), hp(pen), hp(wall), ctypes.byref(vm))
        assert rc == 0
        vel, vsini, params, pen2 = fobj.unpack(idx.astype(np.int64), X)
        wall2 = (vel > 1000) | (vel < -1000) | ~np.isfinite(params).all(axis=1)
        assert wall2.any() and np.array_equal(wall.astype(bool), wall2)
        ok = ~wall2
        assert np.array_equal(h_in[0, :K][ok], vel[ok])
        if vsini is not None:
            assert np.array_equal(h_in[1, :K][ok], vsini[ok]) and vm.value == vsini[ok].max()
        assert np.array_equal(h_in[2:, :K].T[ok], spec_inter.map_params(params, [0])[ok])
        assert np.array_equal(prior[ok], fobj.prior_term(params)[ok])
        assert np.array_equal(pen, pen2)
        assert np.array_equal(h_oix[:, :K][:, ok], eng._oix[:, idx][:, ok])
        assert (h_oix[:, :K][:, ~ok] == -1).all() and (h_oix[:, K:] == -1).all()
        chi = rs.normal(1000, 100, size=(2, narm, Kp))
        chi[1] = 0
        chi[1, :, 5:9] = rs.uniform(0, .1, size=(narm, 4))
        flags = np.zeros((2, narm, Kp), dtype=np.int32)
        flags[0, 1, 7] = 2
        chi[0, 2, 11] = np.nan
        for shared in (0, 1):
            out, redo = np.zeros(K), np.zeros(K, dtype=np.uint8)
            n = L.rvs_fit_collect(ctypes.byref(lay), K, Kp, hp(idx), hp(h_in), hp(chi), hp(flags),
                                  shared, 1, hp(prior), hp(pen), hp(wall), hp(out), hp(redo))
            c, o = chi[0, :, :K], chi[1, :, :K]
            fl = flags[:, :, :K].copy()
            if shared:
                o = np.broadcast_to(o[0], o.shape)
                fl[1, 1:] = fl[1, 0]
            r2 = (fl != 0).any(axis=(0, 1)) | ~np.isfinite(np.stack([c, o])).all(axis=(0, 1))
            r2 |= ~eng._cover0[idx] | (h_in[0, :K] < -1000) | (h_in[0, :K] > 1000)
            present = eng._oix[:, idx] >= 0
            tot = np.add.reduce(np.where(present, o * eng.badchi[idx][None, :], 0.0) + c, axis=0)
            want = prior + tot + pen
            want[wall2] = 1e30
            r2[wall2] = False
            assert n == r2.sum() and np.array_equal(redo.astype(bool), r2)
            fin = np.isfinite(want)
            assert np.array_equal(out[fin], want[fin])


def test_call_size_ladder():
    """rvs_fit_round_items: launch configurations of an evaluation call."""
    from rvspecfit_b200 import _cabi
    L = _cabi.lib()
    seen = set()
    prev = 0
    for K in range(1, 40000):
        Kp = L.rvs_fit_round_items(K)
        assert Kp >= K and Kp % 16 == 0 and Kp >= prev
        assert Kp <= max(16, int(1.25 * K) + 16), (K, Kp)       # at most a quarter of padding
        assert L.rvs_fit_round_items(Kp) == Kp                  # a configuration maps to itself
        prev = Kp
        seen.add(Kp)
    assert len(seen) <= 60          # four per octave: a few dozen captured graphs per slot


def test_stepper_rows_of_stopped_problems_are_final():
    """The hand-over between stages relies on it: while other problems still iterate,
    NMStepper.result() already holds the final row of every problem that has stopped."""
    B, N = 24, 4
    _, fbatch = _problems(B, N, 5)
    rs = np.random.RandomState(11)
    x0 = rs.normal(size=(B, N))
    sims = np.concatenate([x0[:, None, :], x0[:, None, :] + 0.7 * np.eye(N)[None]], axis=1)
    st = batch_fit.NMStepper(sims)
    snapshots = []
    try:
        while True:
            req = st.request(0)
            if req is None:
                break
            idx, X = req
            st.feed(fbatch(idx, X))
            act = st.active()
            if act.any() and not act.all():
                snapshots.append((~act, st.result()))
        final = st.result()
    finally:
        st.close()
    assert len(snapshots) > 5
    for stopped, res in snapshots:
        for k in ('x', 'fun', 'success', 'final_simplex', 'nit', 'nfev'):
            assert np.array_equal(res[k][stopped], final[k][stopped]), k
