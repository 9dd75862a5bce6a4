"""CPU-only checks of the C-ABI boundary: the library loads, exports every
symbol include/rvs_b200.h declares, and the host-side helpers (no device work)
agree with the oracle."""
import ctypes
import os
import re

import numpy as np

from rvspecfit_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, 'include', 'rvs_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(rvs_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    L = _cabi.lib()
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f'{n} declared in rvs_b200.h but not exported'
        assert n in _cabi.SIGNATURES, f'{n} has no ctypes signature'
    assert L.rvs_version() >= 100


def test_struct_layouts_match_header():
    # 5 pointers + 2 int32 + 7 doubles ; 8 pointers + 4 int32 + 2 pointers + 2 int32
    assert ctypes.sizeof(_cabi.Knots) == 5 * 8 + 8 + 7 * 8
    assert ctypes.sizeof(_cabi.Obs) == 8 * 8 + 16 + 2 * 8 + 8
    assert _cabi.Obs.d_resol.offset == 80 and _cabi.Obs.nresol.offset == 96


def test_knot_tables_match_oracle_thomas():
    """cp / winv reproduce the forward elimination of spliner.c:33-41: solving
    with them gives the oracle's spline."""
    import oracle
    L = _cabi.lib()
    x = np.exp(np.linspace(np.log(4000.), np.log(5000.), 300))
    y = np.sin(x / 7.) + 2
    n = len(x)
    h, hinv, cp, winv = np.zeros(n - 1), np.zeros(n - 1), np.zeros(n - 2), np.zeros(n - 2)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    L.rvs_knot_tables(vp(x), n, vp(h), vp(hinv), vp(cp), vp(winv))
    b = (y[1:] - y[:-1]) * hinv
    m = n - 2
    d, z = np.zeros(m), np.zeros(n)
    for k in range(m):
        d[k] = (6 * (b[k + 1] - b[k]) - h[k] * (d[k - 1] if k else 0.0)) * winv[k]
    for k in range(m - 1, -1, -1):
        z[k + 1] = d[k] - cp[k] * z[k + 2]
    s = oracle.Spline(x, y, log_step=True)
    assert np.allclose(z[1:] * hinv / 6, s.A, rtol=1e-12, atol=1e-18)
    kn = _cabi.Knots()
    assert L.rvs_knot_info(vp(x), n, 1, ctypes.byref(kn)) == 0
    assert kn.npix_t == n and np.isclose(kn.lnstep, np.log(x[1] / x[0]))
    xb = x.copy()
    xb[1] *= 1.001
    assert L.rvs_knot_info(vp(xb), n, 1, ctypes.byref(kn)) == -2


def test_compute_path_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        _cabi.require_cuda()
