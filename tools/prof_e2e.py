"""Host-side profile of building a LikelihoodEngine from host arrays (bench e2e leg)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from rvspecfit_b200 import spec_fit, spec_inter
w = bench.WORKLOADS['desi']; cfg = bench.make_config(w)
setups, objects, pars, vel = bench.make_inputs('desi', 2048, 1000)
for st in setups:
    spec_inter.register_bank(spec_inter.bank_from_setup(st), template_lib='synthetic/')
objs = [[spec_fit.SpecData(*a) for a in o] for o in objects]
import torch
for i in range(2):
    t0 = time.time(); eng = spec_fit.LikelihoodEngine(objs, cfg, {'npoly': 10}); torch.cuda.synchronize(); print('build', time.time() - t0)
pr = cProfile.Profile(); pr.enable()
eng = spec_fit.LikelihoodEngine(objs, cfg, {'npoly': 10}); torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
