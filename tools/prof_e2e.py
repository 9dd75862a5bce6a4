"""Host-side profile of one end-to-end bench step (engine from host arrays + hot path)
against the same hot path on a resident engine."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from rvspecfit_b200 import spec_fit, spec_inter, batch_fit


def main():
    import torch
    B, E = 2048, 200
    w = bench.WORKLOADS['desi']; cfg = bench.make_config(w)
    setups, objects, pars, vel = bench.make_inputs('desi', B, 1000)
    for st in setups:
        spec_inter.register_bank(spec_inter.bank_from_setup(st), template_lib='synthetic/')
    objs = [[spec_fit.SpecData(*a) for a in o] for o in objects]
    tp, tv, tvs = bench.trial_points(pars, vel, 'desi', E, 5)
    vgrid = np.arange(cfg['min_vel'], cfg['max_vel'], cfg['vel_step0'])
    start = np.tile(np.array([[5500., 3.0, -1.0, 0.2]]), (B, 1))
    opts = {'npoly': 10}

    def e2e():
        eng = spec_fit.LikelihoodEngine(objs, cfg, opts)
        return batch_fit.scan_and_evaluate(eng, start, vgrid, tp, tv, tvs, groups=2)
    eng0 = spec_fit.LikelihoodEngine(objs, cfg, opts)

    def resident():
        return batch_fit.scan_and_evaluate(eng0, start, vgrid, tp, tv, tvs, groups=2)
    for f in (resident, e2e):
        for _ in range(2):
            f()
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(3):
            f()
        torch.cuda.synchronize(); print(f.__name__, (time.time() - t0) / 3)
    pr = cProfile.Profile(); pr.enable()
    e2e(); torch.cuda.synchronize()
    pr.disable(); pstats.Stats(pr).sort_stats('tottime').print_stats(22)


if __name__ == '__main__':
    main()
