"""Host logic of the batched drivers, no GPU: the lock-step Nelder-Mead and the
threaded BFGS pool must follow scipy's own trajectories bit for bit on an
objective that is cheap to evaluate on the CPU."""
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


def test_threaded_bfgs_pool_is_scipy():
    B, N = 16, 4
    f_one, fbatch = _problems(B, N, 11)
    rs = np.random.RandomState(5)
    x0 = rs.normal(size=(B, N))
    H0 = np.diag([1.0, 25.0, 0.01, 4.0])
    calls = []

    def counted(idx, X):
        calls.append(len(idx))
        return fbatch(idx, X)
    got = batch_fit.bfgs_batch(counted, x0, H0)
    for b in range(B):
        want = scipy.optimize.minimize(lambda x: f_one(b, x), x0[b], method='BFGS',
                                       options=dict(hess_inv0=H0))
        assert np.array_equal(got[b]['x'], want['x']), b
        assert got[b]['fun'] == want['fun'] and got[b]['nit'] == want['nit']
    # requests were gathered: far fewer batched calls than scalar evaluations
    assert len(calls) < sum(calls) / 4
    # the same through worker processes
    got2 = batch_fit.bfgs_many(fbatch, x0, H0, nproc=2)
    for b in range(B):
        assert np.array_equal(got2[b]['x'], got[b]['x']) and got2[b]['nit'] == got[b]['nit']


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
