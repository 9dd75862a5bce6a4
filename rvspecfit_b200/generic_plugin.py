"""GPU-backed interpolator for the reference's own plugin hook.

Reference spec_inter.getInterpolator (spec_inter.py:371-378): an
`interp_<setup>.h5` whose `interpolation_type` is `'generic'` names a module and
two classes; the reference imports the module, instantiates both classes with
the loaded dictionary (plus `template_lib`) and then calls

    interper(mapped_params) -> spectrum  f64[Npix_t]    (SpecInterpolator.eval)
    extraper(mapped_params) -> off-grid measure, 0 inside (SpecInterpolator.outsideFlag)

with the parameter vector ALREADY passed through the product's mapper
(spec_inter.py:268-285).  `GpuGridInterp` / `GpuGridOutside` are such classes:
an UNMODIFIED rvspecfit then evaluates its templates with rvs_template_build on
the device (grid rows resident in HBM) instead of GridInterp's fancy-index gather
+ dot + exp (spec_inter.py:134-194), and gets GridOutsideCheck's measure
(spec_inter.py:77-92) from the same bank.  This is the per-call drop-in; the
throughput path is the batched likelihood (spec_fit.LikelihoodEngine).

The dictionary is a regular-grid product (`uvecs`, `idgrid`, `vec`, `lam`,
`parnames`, `log_step`, optionally `log_spec`) with the three plugin keys added by
`generic_fd`; the grid rows come from `dats` in the dictionary or from
`interpdat_<setup>.npy` under `template_lib` (make_nd.py:14-15).
"""
import os

import numpy as np

MODULE = 'rvspecfit_b200.generic_plugin'
_banks = {}


def generic_fd(fd, setup):
    """Turn a loaded regular-grid product dictionary into one that selects this
    plugin (what a maintainer writes back with serializer.save_dict_to_hdf5)."""
    out = dict(fd)
    out.update(interpolation_type='generic', module=MODULE, class_name='GpuGridInterp',
               outside_class_name='GpuGridOutside', setup=setup)
    return out


def _bank(fd):
    from . import spec_inter
    key = (fd.get('template_lib'), fd['setup'])
    if key not in _banks:
        dats = fd.get('dats')
        if dats is None:
            dats = np.load(os.path.join(fd['template_lib'], 'interpdat_%s.npy' % fd['setup']),
                           mmap_mode='r')
        # log_ids=(): the reference has applied the product's mapper before calling us
        _banks[key] = spec_inter.TemplateBank(
            fd['setup'], fd['lam'], dats, [str(_) for _ in fd['parnames']], kind='regulargrid',
            uvecs=fd['uvecs'], idgrid=fd['idgrid'], vecs=fd['vec'], log_ids=(),
            log_step=bool(fd['log_step']), log_spec=bool(fd.get('log_spec', True)))
    return _banks[key]


class GpuGridInterp:
    """`class_name` of the generic hook: polylinear template evaluation."""

    def __init__(self, fd):
        self.bank = _bank(fd)

    def __call__(self, p):
        spec, _ = self.bank.template(np.asarray(p, dtype=np.float64)[None, :])
        return spec[0]


class GpuGridOutside:
    """`outside_class_name` of the generic hook: 0 inside the grid, else the distance
    to the nearest node in peak-to-peak normalised coordinates."""

    def __init__(self, fd):
        self.bank = _bank(fd)

    def __call__(self, p):
        return float(self.bank.locate(np.asarray(p, dtype=np.float64)[None, :])[2][0])
