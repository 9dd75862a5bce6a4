"""Maximum-likelihood fit of one object (and the brute-force first guess).

Mirror of the reference's vel_fit.py (paths under /root/reference/py/rvspecfit/):
`process` and `firstguess` keep their signatures, returned keys and decision
rules -- the same scipy optimisers with the same options drive the same
objective, whose every evaluation is a GPU call (spec_fit.LikelihoodEngine).
`process_batch` (batch_fit.py) is the throughput path that steps many objects
in lock-step.
"""
import copy
import itertools
import logging
import math

import numpy as np
import scipy.linalg
import scipy.optimize

from . import spec_fit, spec_inter


def firstguess(specdata, options=None, config=None, resolParams=None,
               vsinigrid=(None, 10, 100), paramsgrid=None):
    """Brute-force starting point over a small template grid x vsini x RV grid
    (reference vel_fit.py:13-94)."""
    min_vel, max_vel, vel_step0 = config['min_vel'], config['max_vel'], config['vel_step0']
    options = options or {}
    if paramsgrid is None:
        paramsgrid = {'logg': [1, 2, 3, 4, 5], 'teff': [3000, 5000, 8000, 10000],
                      'feh': [-2, -1, 0], 'alpha': [0]}
    if isinstance(specdata, spec_fit.SpecData):
        specdata = [specdata]
    specParams = spec_inter.getSpecParams(specdata[0].name, config)
    params = []
    for x in itertools.product(*paramsgrid.values()):
        curp = dict(zip(paramsgrid.keys(), x))
        params.append([curp[_] for _ in specParams])
    vels_grid = np.arange(min_vel, max_vel, vel_step0)
    best_chisq = np.inf
    for vsini in vsinigrid:
        rot_params = None if vsini is None else (vsini, )
        res = spec_fit.find_best(specdata, vels_grid, params, rot_params=rot_params,
                                 resol_params=resolParams, config=config, options=options)
        if res['best_chi'] < best_chisq:
            bestpar = {k: res['best_param'][i] for i, k in enumerate(specParams)}
            if vsini is not None:
                bestpar['vsini'] = vsini
            best_chisq = res['best_chi']
    return bestpar


class VSiniMapper:
    """reference vel_fit.py:97-116."""

    def __init__(self, max_vsini):
        self.max_vsini = max_vsini

    def to_internal(self, vsini):
        return np.clip(vsini, 0, self.max_vsini)

    def to_vsini(self, x):
        vsini = np.clip(x, 0, self.max_vsini)
        penalty = int(x < 0) * (vsini - x)**2 + int(x > self.max_vsini) * (vsini - x)**2
        return vsini, penalty


class ParamMapper:
    """Fitted vector <-> named parameters; vector layout
    [vel, (vsini), free atmospheric parameters] (reference vel_fit.py:119-207)."""

    def __init__(self, specParams, paramDict0, fixParam, vsiniMapper, fitVsini=True):
        self.specParams, self.paramDict0, self.fixParam = specParams, paramDict0, fixParam
        self.vsiniMapper, self.fitVsini = vsiniMapper, fitVsini

    def forward(self, p0):
        ret = {}
        rev = list(p0)[::-1]
        penalty = 0
        ret['vel'] = rev.pop()
        if self.fitVsini:
            vsini, pen = self.vsiniMapper.to_vsini(rev.pop())
            penalty += pen
            ret['vsini'] = vsini
        else:
            ret['vsini'] = self.paramDict0['vsini'] if 'vsini' in self.fixParam else None
        ret['rot_params'] = None if ret['vsini'] is None else (ret['vsini'], )
        ret['params'] = [self.paramDict0[x] if x in self.fixParam else rev.pop()
                         for x in self.specParams]
        assert len(rev) == 0
        ret['penalty'] = penalty
        return ret

    def get_fitted_params(self):
        return ['vel'] + (['vsini'] if self.fitVsini else []) + \
            [x for x in self.specParams if x not in self.fixParam]


def chisq_func0(pdict, args, outside_penalty=True):
    """-2 log L + priors (reference vel_fit.py:210-230)."""
    chisq = 0
    if args.get('priors') is not None:
        priors = args['priors']
        for i, k in enumerate(args['paramMapper'].specParams):
            if k in priors:
                chisq += ((priors[k][0] - pdict['params'][i]) / priors[k][1])**2
    chisq += spec_fit.get_chisq(args['specdata'], pdict['vel'], pdict['params'],
                                pdict['rot_params'], args['resolParams'],
                                options=args['options'], config=args['config'],
                                outside_penalty=outside_penalty)
    return chisq


def chisq_func(p, args):
    """Objective of the optimisers (reference vel_fit.py:233-257)."""
    pdict = args['paramMapper'].forward(p)
    if (pdict['vel'] > args['max_vel'] or pdict['vel'] < args['min_vel']
            or (~np.isfinite(pdict['params'])).any()):
        return 1e30
    return chisq_func0(pdict, args) + pdict['penalty']


def hess_func(p, pdict, args):
    """reference vel_fit.py:260-269."""
    pdict['params'][:] = p[:]
    return 0.5 * chisq_func0(pdict, args)


SIMPLEX_STD = {'logg': 0.5, 'teff': 300, 'feh': 0.5, 'alpha': 0.25}


def _get_simplex_start(best_vel, fixParam=None, specParamNames=None, paramDict0=None,
                       vsiniMapper=None, fitVsini=None):
    """Deterministic starting simplex (reference vel_fit.py:272-312)."""
    startParam, std_vec = [best_vel], [5]
    if fitVsini:
        startParam.append(vsiniMapper.to_internal(paramDict0['vsini']))
        std_vec.append(3)
    for x in specParamNames:
        if x not in fixParam:
            startParam.append(paramDict0[x])
            std_vec.append(SIMPLEX_STD.get(x) or 0.5)
    curval, std_vec = np.array(startParam), np.array(std_vec)
    ndim = len(curval)
    R = np.random.RandomState(43434)
    simp = np.zeros((ndim + 1, ndim))
    simp[0, :] = curval
    simp[1:, :] = curval[None, :] + std_vec[None, :] * R.normal(size=(ndim, ndim))
    return curval, simp


def _minimum_sampler(func, best_vel, min_vel, max_vel, vel_step0, min_vel_step,
                     crit_ratio=5, goal_width=10):
    """Shrinking-grid refinement of the RV posterior (reference vel_fit.py:358-439)."""
    vel_step = vel_step0
    for it in range(10):
        vels_grid = np.arange(math.ceil((min_vel - best_vel) / vel_step) * vel_step,
                              max_vel - best_vel, vel_step) + best_vel
        best_vel, cur_err, res1 = func(vels_grid)
        if vel_step < cur_err / crit_ratio or vel_step < min_vel_step:
            break
        if vel_step > cur_err:
            vel_step_new, width_new = vel_step / crit_ratio, vel_step * goal_width
        else:
            vel_step_new, width_new = cur_err / crit_ratio * 0.8, cur_err * goal_width
        min_vel = max(best_vel - width_new, min_vel)
        max_vel = min(best_vel + width_new, max_vel)
        vel_step = vel_step_new
    if it > 5:
        logging.warning('More than 5 iterations we used in finding the velocity error')
    return best_vel, cur_err, res1


def _find_best_vel_iterate(best_vel, min_vel, max_vel, vel_step0, specdata=None,
                           best_param=None, resolParams=None, config=None, options=None,
                           min_vel_step=None):
    """reference vel_fit.py:315-355."""
    if best_vel > max_vel or best_vel < min_vel:
        logging.warning('Velocity too large...')
        best_vel = max_vel if best_vel > max_vel else min_vel

    def func(vels_grid):
        res1 = spec_fit.find_best(specdata, vels_grid, [best_param['params']],
                                  rot_params=best_param['rot_params'],
                                  resol_params=resolParams, config=config, options=options)
        return res1['best_vel'], res1['vel_err'], res1
    best_vel, best_err, res1 = _minimum_sampler(func, best_vel, min_vel, max_vel, vel_step0,
                                                min_vel_step)
    return best_vel, best_err, res1['skewness'], res1['kurtosis']


def get_hess_inv(param_names):
    """Initial inverse Hessian of BFGS (reference vel_fit.py:442-460)."""
    diag = np.zeros(len(param_names)) + 0.1**2
    diag[np.nonzero(np.asarray(param_names) == 'teff')[0][0]] = 50**2
    vs = np.nonzero(np.asarray(param_names) == 'vsini')[0]
    if len(vs) == 1:
        diag[vs] = 5**2
    diag[0] = 1
    return np.diag(diag)


def _uncertainties_from_hessian(hessian):
    """reference vel_fit.py:463-502."""
    diag_hessian = np.diag(hessian)
    inv_diag = 1. / (diag_hessian + (diag_hessian == 0))
    inv_diag[diag_hessian == 0] = np.inf
    bad_hessian = False
    try:
        hessian_inv = scipy.linalg.inv(hessian)
    except (np.linalg.LinAlgError, ValueError):
        bad_hessian = True
        logging.warning('The inversion of the Hessian failed')
        hessian_inv = np.diag(inv_diag)
    diag_err0 = np.array(np.diag(hessian_inv))
    bad0, bad1 = diag_err0 < 0, inv_diag < 0
    if bad0.any():
        bad_hessian = True
    sub1, sub2 = bad0 & (~bad1), bad0 & bad1
    diag_err0[sub1] = inv_diag[sub1]
    diag_err0[sub2] = 0
    diag_err = np.sqrt(diag_err0)
    diag_err[sub2] = np.nan
    if (~np.isfinite(diag_err)).sum() != 0:
        bad_hessian = True
    return diag_err, hessian_inv, bad_hessian


HESS_STEP = {'vsini': 1 / 100, 'logg': 0.1 / 100, 'feh': 0.1 / 100, 'alpha': .01 / 100,
             'teff': 1 / 100, 'vrad': 1 / 100}   # reference vel_fit.py:705-712


def central_hessian(f, x, steps):
    """Central-difference Hessian with one Richardson step (h, h/2).

    The reference calls numdifftools.Hessian here (vel_fit.py:713-716);
    numdifftools is not available offline, so when it cannot be imported this
    routine is used instead (parity of param_err is unpinned)."""
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


def _hessian(func, x, steps, use_steps):
    try:
        import numdifftools as ndf
        gen = ndf.MinStepGenerator(base_step=steps) if use_steps else None
        return ndf.Hessian(func, step=gen)(x)
    except ImportError:
        return central_hessian(func, x, steps)


def process(specdata, paramDict0, fixParam=None, options=None, config=None,
            resolParams=None, priors=None):
    """Maximum-likelihood fit of one object: reference vel_fit.py:505-737, same
    arguments, same returned keys."""
    if config is None:
        raise RuntimeError('Config must be provided')
    if isinstance(specdata, spec_fit.SpecData):
        specdata = [specdata]
    min_vel, max_vel = config['min_vel'], config['max_vel']
    vel_step0, max_vsini = config['vel_step0'], config['max_vsini']
    min_vel_step = config['min_vel_step']
    second_minimizer = config.get('second_minimizer') or False
    options = options or {}
    vels_grid = np.arange(min_vel, max_vel, vel_step0)
    curparam = spec_fit.param_dict_to_tuple(paramDict0, specdata[0].name, config=config)
    specParamNames = spec_inter.getSpecParams(specdata[0].name, config)
    if fixParam is None:
        fixParam = []
    vsiniMapper = None
    if 'vsini' not in paramDict0:
        rot_params, fitVsini = None, False
    else:
        rot_params = (paramDict0['vsini'], )
        fitVsini = 'vsini' not in fixParam
        if fitVsini:
            vsiniMapper = VSiniMapper(max_vsini)
    res = spec_fit.find_best(specdata, vels_grid, [curparam], rot_params=rot_params,
                             resol_params=resolParams, config=config, options=options)
    best_vel = res['best_vel']
    curval, simplex = _get_simplex_start(best_vel, fixParam=fixParam,
                                         specParamNames=specParamNames,
                                         paramDict0=paramDict0, vsiniMapper=vsiniMapper,
                                         fitVsini=fitVsini)
    paramMapper = ParamMapper(specParamNames, paramDict0, fixParam, vsiniMapper,
                              fitVsini=fitVsini)
    args = dict(min_vel=min_vel, max_vel=max_vel, resolParams=resolParams,
                paramMapper=paramMapper, specdata=specdata, options=options, config=config,
                priors=priors)
    minimize_success = True
    curiter, maxiter = 1, 2
    hess_inv0 = get_hess_inv(paramMapper.get_fitted_params())
    while True:
        res0 = scipy.optimize.minimize(chisq_func, curval, args=args, method='Nelder-Mead',
                                       options={'fatol': 1e-3, 'xatol': 1e-2,
                                                'initial_simplex': simplex,
                                                'maxiter': 10000, 'maxfev': np.inf})
        curval = res0['x']
        simplex = res0['final_simplex'][0]
        if res0['success']:
            break
        if curiter == maxiter:
            logging.warning('Maximum number of iterations reached')
            minimize_success = False
            break
        curiter += 1
    if second_minimizer:
        res = scipy.optimize.minimize(chisq_func, res0['x'], method='BFGS', args=args,
                                      options=dict(hess_inv0=hess_inv0))
    else:
        res = res0
    best_param = paramMapper.forward(res['x'])
    ret = {}
    ret['param'] = dict(zip(specParamNames, best_param['params']))
    if fitVsini:
        ret['vsini'] = best_param['vsini']
    best_vel = best_param['vel']
    best_vel, vel_err, vel_skewness, vel_kurtosis = _find_best_vel_iterate(
        best_vel, min_vel, max_vel, vel_step0, specdata=specdata, best_param=best_param,
        resolParams=resolParams, config=config, options=options, min_vel_step=min_vel_step)
    ret['vel'], ret['vel_err'] = best_vel, vel_err
    ret['vel_skewness'], ret['vel_kurtosis'] = vel_skewness, vel_kurtosis
    outp = spec_fit.get_chisq(specdata, best_vel, best_param['params'],
                              best_param['rot_params'], resolParams, options=options,
                              config=config, full_output=True)
    best_param_TMP = copy.deepcopy(best_param)

    def hess_func_wrap(p):
        return hess_func(p, best_param_TMP, args)
    hess_step = [HESS_STEP[_] for _ in specParamNames]
    use_steps = True
    for i in range(2):
        hessian = _hessian(hess_func_wrap, [ret['param'][_] for _ in specParamNames],
                           hess_step, use_steps)
        diag_err, covar_mat, bad_hessian = _uncertainties_from_hessian(hessian)
        if bad_hessian:
            use_steps = False
            logging.warning('Performing two iterations of hessian determination')
    ret['param_err'] = dict(zip(specParamNames, diag_err))
    ret['param_covar'] = covar_mat
    ret['minimize_success'] = minimize_success
    ret['bad_hessian'] = bad_hessian
    ret['yfit'] = outp['models']
    ret['raw_models'] = outp['raw_models']
    ret['chisq'] = outp['chisq']
    ret['logl'] = outp['logl']
    ret['chisq_array'] = outp['chisq_array']
    ret['npix_array'] = outp['npix_array']
    return ret
