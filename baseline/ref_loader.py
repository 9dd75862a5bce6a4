"""Import the reference package installed by baseline/install_ref.py, with the absent
third-party modules stubbed: h5py / astropy / matplotlib are only
touched by I/O and plotting code that the hot path never reaches; numdifftools' Hessian
(vel_fit.py:713-716) is replaced by the central-difference routine that the oracle and the
product use too (param_err parity is unpinned).  Used by bench.py's reference arm
and by tests/golden/make_golden.py -- never by the product."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')

HESS_STEP = {'vsini': 1 / 100, 'logg': 0.1 / 100, 'feh': 0.1 / 100, 'alpha': .01 / 100,
             'teff': 1 / 100, 'vrad': 1 / 100}    # vel_fit.py:705-712


def available():
    return os.path.exists(os.path.join(REF_DIR, 'rvspecfit', '_version.py'))


def load(parnames=('teff', 'logg', 'feh', 'alpha')):
    """Namespace of the reference's modules on the hot path."""
    if not available():
        raise RuntimeError('reference not installed: run python baseline/install_ref.py in the '
                           'build container')
    for m in ['h5py', 'astropy', 'astropy.io', 'astropy.io.fits', 'numdifftools',
              'matplotlib', 'matplotlib.pyplot']:
        sys.modules.setdefault(m, types.ModuleType(m))
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import oracle
    ndf = sys.modules['numdifftools']

    class MinStepGenerator:
        def __init__(self, base_step=None):
            self.base_step = base_step

    class Hessian:
        def __init__(self, f, step=None):
            self.f, self.step = f, step

        def __call__(self, x):
            steps = self.step.base_step if self.step is not None else \
                [HESS_STEP[k] for k in parnames]
            return oracle.central_hessian(self.f, x, steps)
    ndf.MinStepGenerator, ndf.Hessian = MinStepGenerator, Hessian
    import rvspecfit  # noqa: F401
    from rvspecfit import (spec_fit, spec_inter, vel_fit, fitter_ccf, make_ccf, read_grid,
                           utils, spliner, make_nd)
    return types.SimpleNamespace(spec_fit=spec_fit, spec_inter=spec_inter, vel_fit=vel_fit,
                                 fitter_ccf=fitter_ccf, make_ccf=make_ccf, read_grid=read_grid,
                                 utils=utils, spliner=spliner, make_nd=make_nd, oracle=oracle)


def inject_grid(R, setup, name=None):
    """Register a synthetic regular-grid bank (rvspecfit_b200/synth.py) with the reference's
    interpolator cache, bypassing its HDF5 loader."""
    name = name or setup['name']
    si = R.spec_inter
    si.interp_cache.template_lib = 'synthetic/'
    it = si.SpecInterpolator(
        name, si.GridInterp(setup['uvecs'], setup['idgrid'], setup['vec'], setup['dats'],
                            exp=True),
        si.GridOutsideCheck(setup['uvecs'], setup['vec'], setup['idgrid']),
        setup['lam'], R.read_grid.LogParamMapper([0]), setup['parnames'], log_step=True)
    si.interp_cache.interps[name] = it
    return it
