"""Host-side profile of batch_fit.process_batch (bench --mode fit) on a DESI-shaped batch."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from rvspecfit_b200 import spec_fit, spec_inter, batch_fit
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512


def main():
    w = bench.WORKLOADS['desi']; cfg = bench.make_config(w)
    setups, objects, pars, vel = bench.make_inputs('desi', B, 1000)
    for st in setups:
        spec_inter.register_bank(spec_inter.bank_from_setup(st), template_lib='synthetic/')
    objs = [[spec_fit.SpecData(*a) for a in o] for o in objects]
    eng = spec_fit.LikelihoodEngine(objs, cfg, {'npoly': 10})
    starts = [dict(bench.FIT_START) for _ in range(B)]
    batch_fit.process_batch(None, starts, config=cfg, options={'npoly': 10}, engine=eng)
    pr = cProfile.Profile(); pr.enable(); t0 = time.time()
    batch_fit.process_batch(None, starts, config=cfg, options={'npoly': 10}, engine=eng)
    pr.disable(); print('wall', time.time() - t0, batch_fit.process_batch.last_phase_seconds)
    pstats.Stats(pr).sort_stats('tottime').print_stats(28)


if __name__ == '__main__':
    main()
