"""Shared helpers for the parity tests: rebuild the inputs the golden fixtures
were made from (tests/golden/make_golden.py) and unpack stored objects."""
import numpy as np

from rvspecfit_b200 import synth

CONFIG = dict(min_vel=-1000, max_vel=1000, vel_step0=5, max_vsini=500, min_vsini=0.1,
              min_vel_step=0.2, second_minimizer=True, template_lib='synthetic/')


# parameter vectors probing interior / edge / off-grid cases of the 'tiny' layout
PROBE_PARAMS = [
    (5000., 2.0, -1.0, 0.2),       # interior
    (4333., 3.71, -0.42, 0.66),    # interior, generic
    (3500., 2.5, -1.0, 0.5),       # exactly on lowest teff node of 'tiny'
    (9000., 2.5, -1.0, 0.5),       # top edge: counts as outside
    (5000., 5.0, -1.0, 0.5),       # top edge in logg
    (2000., 2.0, -1.0, 0.2),       # below grid
    (20000., 6.0, 1.0, 2.0),       # far outside
    (5000., 2.0, -1.0, -0.3),      # outside in alpha
    (6100., 1.3, -1.9, 0.05),      # interior near corner
]


def config(**kw):
    c = dict(CONFIG)
    c.update(kw)
    return c


def unpack_objects(g, prefix):
    objs = []
    for i in range(int(g[prefix + 'n'])):
        names = [str(_) for _ in g[f'{prefix}{i}_names']]
        arms = [(nm, g[f'{prefix}{i}_{a}_lam'], g[f'{prefix}{i}_{a}_spec'],
                 g[f'{prefix}{i}_{a}_espec'], g[f'{prefix}{i}_{a}_bad'])
                for a, nm in enumerate(names)]
        objs.append(dict(params=g[prefix + 'params'][i], vel=float(g[prefix + 'vel'][i]),
                         arms=arms))
    return objs


_setups = {}


def setup(shape, layout, seed, holes=0, name=None):
    key = (shape, layout, seed, holes)
    if key not in _setups:
        _setups[key] = synth.make_setup(shape, layout, seed=seed, holes=holes)
    st = dict(_setups[key])
    if name is not None:
        st['name'] = name
    return st


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def close(a, b, rtol, atol=0.0, what=''):
    """allclose that reports the worst deviation in units of the tolerance."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f'{what}: NaN pattern differs'
    err = np.abs(a - b)[~nan_a]
    tol = (atol + rtol * np.abs(b))[~nan_a]
    if err.size == 0:
        return True
    worst = np.max(err / np.maximum(tol, 1e-300))
    assert worst <= 1, f'{what}: worst deviation {worst:.3g} x tolerance (rtol={rtol}, atol={atol})'
    return True
