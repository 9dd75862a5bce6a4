import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import oracle
from helpers import setup, unpack_objects, config
from rvspecfit_b200 import spec_fit, spec_inter, _dev, _cabi
g = dict(np.load(os.path.join(ROOT, 'tests/golden/chisq.npz')))
st = setup('test', 'tiny', 3, name='test')
bank = spec_inter.bank_from_setup(st)
spec_inter.register_bank(bank, 'synthetic/')
objs = unpack_objects(g, 'one_'); ev = g['one_eval']; cfg = config()
s1 = spec_fit.SpecData(*objs[0]['arms'][0])
dbg = _dev.zeros((8192,), np.float64)
_cabi.lib().rvs_set_debug_buffer(_dev.ptr(dbg))
eng = spec_fit.LikelihoodEngine([[s1]], cfg, {'npoly': 15})
e = ev[0]
c1 = eng.evaluate([0], np.array([e[0]]), e[None, 1:5], None)
d = _dev.download(dbg)
ya0, ya1, kr0, kr1, c0, c1_, plo, phi, g0a, W0 = [int(_) for _ in d[:10]]
print('ya0,ya1,kr0,kr1,c0,c1,plo,phi,g0a,W0', ya0, ya1, kr0, kr1, c0, c1_, plo, phi, g0a, W0)
wcap = (936 + 2 * 35 + 20 + 3) & ~3
Y = d[16:16 + ya1 - ya0 + 1]; Z = d[16 + wcap:16 + wcap + kr1 - kr0]
spec, _ = bank.template(e[None, 1:5])
ytrue = spec[0]
print('y err', np.abs(Y - ytrue[ya0:ya1 + 1]).max())
sp = oracle.Spline(st['lam'], np.ascontiguousarray(ytrue))
ztrue = np.concatenate([[0], sp.A * 6 * sp.h])
zt = ztrue[kr0 + 1:kr1 + 1]
err = np.abs(Z - zt)
print('z err max', err.max(), 'at', err.argmax(), 'of', len(Z), 'zmax', np.abs(zt).max())
print('z err head', err[:40]); print('z err mid', err[400:410]); print('tail', err[-40:])
tn = _dev.download(eng._ws[:len(s1.lam)])
c2, info = eng.evaluate([0], np.array([e[0]]), e[None, 1:5], None, want_model=True)
raw = info['arms']['test']['extras']['raw']
Tf = tn * s1.espec
print('fused T vs general raw: max rel err', np.abs(Tf - raw).max() / np.abs(raw).max())
# host evaluation from the dumped window
lam_t = st['lam']; n = len(lam_t)
beta = e[0] / 299792.458; f = np.sqrt((1 - beta) / (1 + beta))
x = s1.lam * f
q0 = np.log(lam_t[0]); qinv = 1 / np.log(lam_t[1] / lam_t[0])
pos = ((np.log(s1.lam) + np.log(f) - q0) * qinv).astype(int)
h = np.diff(lam_t); hinv = 1 / h
y0 = Y[pos - ya0]; y1 = Y[pos + 1 - ya0]; z0 = Z[pos - 1 - kr0]; z1 = Z[pos - kr0]
t1 = hinv[pos] / 6; t2 = h[pos] / 6
A = z1 * t1; B = z0 * t1; C = y1 * hinv[pos] - z1 * t2; D = y0 * hinv[pos] - z0 * t2
dl = x - lam_t[pos]; dr = lam_t[pos + 1] - x
Th = A * dl**3 + B * dr**3 + C * dl + D * dr
print('host-from-dump vs general raw', np.abs(Th - raw).max() / np.abs(raw).max())
print('host-from-dump vs fused', np.abs(Th - Tf).max() / np.abs(raw).max())
print('pos range', pos.min(), pos.max(), 'chisq fused', c1, 'general', c2)
bad = np.nonzero(np.abs(Tf - raw) > 1e-9)[0]
print('bad px', len(bad), bad[:10], (Tf - raw)[bad[:10]])
dd = d[16 + 2 * wcap:16 + 2 * wcap + 8 * 16].reshape(8, 16)
for p_ in range(4):
    o = dd[p_]
    print('px', p_, 'dev pos', o[0], 'host pos', pos[p_], 'x', o[1] - x[p_], 'y0', o[2] - y0[p_], 'y1', o[3] - y1[p_],
          'z0', o[4] - z0[p_], 'z1', o[5] - z1[p_], 'xl', o[6] - lam_t[pos[p_]], 'xr', o[7] - lam_t[pos[p_] + 1],
          'h', o[8] - h[pos[p_]], 'hinv', o[9] - hinv[pos[p_]], 'einv', o[10] - 1 / s1.espec[p_], 'tn', o[11] - Th[p_] / s1.espec[p_],
          'lam', o[12] - s1.lam[p_], 'f', o[13] - f)
for p_ in range(4):
    o = dd[p_]
    i5 = np.argmin(np.abs(Z - o[5])); i4 = np.argmin(np.abs(Z - o[4]))
    print('px', p_, 'pos-kr0', pos[p_] - kr0, 'z0v found at slot', i4, 'z1v found at slot', i5, 'resid', np.abs(Z - o[5]).min())
