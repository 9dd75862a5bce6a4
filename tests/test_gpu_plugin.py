"""The reference's `generic` interpolator hook (spec_inter.py:371-378) served by the
GPU bank: the classes are resolved and called exactly as the reference does it, and
return what the reference's own GridInterp / GridOutsideCheck returned for the same
product (tests/golden/interp.npz).  Needs a B200."""
import importlib

import numpy as np
import pytest

from helpers import PROBE_PARAMS, close, setup
from rvspecfit_b200 import generic_plugin, spec_inter

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('tag,holes', [('grid', 0), ('holes', 3)])
def test_generic_hook_returns_reference_values(golden, tag, holes):
    g = golden('interp')
    st = setup('test', 'tiny', 3, holes=holes)
    product = dict(lam=st['lam'], parnames=list(st['parnames']), uvecs=st['uvecs'],
                   idgrid=st['idgrid'], vec=st['vec'], log_step=True, dats=st['dats'],
                   mapper_module='rvspecfit.read_grid', mapper_class_name='LogParamMapper',
                   mapper_args=([0],))
    fd = generic_plugin.generic_fd(product, 'plug_' + tag)
    assert fd['interpolation_type'] == 'generic'
    # --- the reference's own lines (spec_inter.py:371-378)
    mod = importlib.import_module(fd['module'])
    fd['template_lib'] = 'synthetic/'
    interper = getattr(mod, fd['class_name'])(fd)
    extraper = getattr(mod, fd['outside_class_name'])(fd)
    # --- SpecInterpolator.eval / outsideFlag: mapper.forward, then the two callables
    for i, p in enumerate(PROBE_PARAMS):
        q = spec_inter.map_params(p, (0,))[0]
        spec, out = interper(q), extraper(q)
        want, wout = g[f'{tag}_spec'][i], g[f'{tag}_outside'][i]
        assert spec.shape == want.shape and spec.dtype == np.float64
        close(out, wout, rtol=1e-13)
        # off-grid: nearest node, exp of the float32 row (<= 1 fp32 ulp, see
        # test_template_interpolation)
        close(spec, want, rtol=1e-13 if wout == 0 else 4e-7)
    # both callables share one bank in HBM
    assert interper.bank is extraper.bank
