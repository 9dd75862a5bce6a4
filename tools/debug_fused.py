import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import setup, unpack_objects, config
from rvspecfit_b200 import spec_fit, spec_inter, _dev
g = dict(np.load(os.path.join(ROOT, 'tests/golden/chisq.npz')))
which = sys.argv[1] if len(sys.argv) > 1 else 'test'
if which == 'test':
    spec_inter.register_bank(spec_inter.bank_from_setup(setup('test', 'tiny', 3, name='test')), 'synthetic/')
    objs = unpack_objects(g, 'one_'); ev = g['one_eval']; cfg = config(); npoly = 15
else:
    for k, a in enumerate(('desi_b', 'desi_r', 'desi_z')):
        spec_inter.register_bank(spec_inter.bank_from_setup(setup(a, 'tiny', 21 + k)), 'synthetic/')
    objs = unpack_objects(g, 'desi_'); ev = g['desi_eval']; cfg = config(min_vel=-1500, max_vel=1500); npoly = 10
sd = [spec_fit.SpecData(*a) for a in objs[0]['arms']]
for e in ev[:4]:
    vs = None if e[5] < 0 else np.array([e[5]])
    for arm_i, s1 in enumerate(sd):
        eng = spec_fit.LikelihoodEngine([[s1]], cfg, {'npoly': npoly})
        c1 = eng.evaluate([0], np.array([e[0]]), e[None, 1:5], vs)
        tn = _dev.download(eng._ws[:len(s1.lam)])
        c2, info = eng.evaluate([0], np.array([e[0]]), e[None, 1:5], vs, want_model=True)
        raw = info['arms'][s1.name]['extras']['raw']
        ref = raw / s1.espec
        err = np.abs(tn - ref) / np.abs(ref).max()
        bad = np.nonzero(err > 1e-12)[0]
        print(s1.name, 'vsini', e[5], 'chisq fused', c1[0], 'general', c2[0], 'max tn err', err.max(),
              'nbad', len(bad), 'first/last bad', (bad[:3], bad[-3:]) if len(bad) else None)
