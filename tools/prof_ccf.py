"""Host-side profile of fitter_ccf.fit_batch on the Gaia-RVS CCF workload of bench.py."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from rvspecfit_b200 import spec_fit, spec_inter, fitter_ccf
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048


def main():
    import torch
    w = bench.WORKLOADS['gaia_rvs']
    cfg = bench.make_config(w)
    setups, objects, pars, vel = bench.make_inputs('gaia_rvs', B, 1000)
    for st in setups:
        spec_inter.register_bank(spec_inter.bank_from_setup(st), template_lib='synthetic/')
    conf, models, params, vsinis = bench.ccf_bank('gaia_rvs', setups[0])
    name = setups[0]['name']
    fitter_ccf.register_ccf_bank(name, np.fft.rfft(models, axis=1), np.fft.rfft(models**2, axis=1),
                                 models, params, vsinis, list(setups[0]['parnames']), conf)
    sds = [[spec_fit.SpecData(*a) for a in o] for o in objects]
    for _ in range(2):
        fitter_ccf.fit_batch(sds, cfg, want_proc_spec=False)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.time()
    pr.enable()
    fitter_ccf.fit_batch(sds, cfg, want_proc_spec=False)
    torch.cuda.synchronize()
    pr.disable()
    print('wall', time.time() - t0)
    pstats.Stats(pr).sort_stats('tottime').print_stats(22)


if __name__ == '__main__':
    main()
