"""Tuning sweep of batch_fit.process_batch on one DESI-shaped batch (engine built once):
python tools/tune_fit.py B "groups:speculate_below[:nm_min]" ...   -> fits/s per setting."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from rvspecfit_b200 import spec_fit, spec_inter, batch_fit, _dev

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
settings = sys.argv[2:] or ['4:128']


def main():
    import torch
    if os.environ.get('RVS_SWITCH'):
        sys.setswitchinterval(float(os.environ['RVS_SWITCH']))
    w = bench.WORKLOADS['desi']
    cfg = bench.make_config(w)
    setups, objects, pars, vel = bench.make_inputs('desi', B, 1000)
    for st in setups:
        spec_inter.register_bank(spec_inter.bank_from_setup(st), template_lib='synthetic/')
    objs = [[spec_fit.SpecData(*a) for a in o] for o in objects]
    eng = spec_fit.LikelihoodEngine(objs, cfg, {'npoly': 10})
    starts = [dict(bench.FIT_START) for _ in range(B)]
    first = True
    for s in settings:
        f = s.split(':')
        groups, spec = int(f[0]), int(f[1])
        threads = bool(int(f[2])) if len(f) > 2 and f[2] != '' else None
        peel = bool(int(f[3])) if len(f) > 3 else None
        if len(f) > 4:
            batch_fit.PEEL_MIN = int(f[4])
        if len(f) > 5:
            batch_fit.PEEL_FRAC = float(f[5])
        if len(f) > 6:
            batch_fit.FIT_SPLIT = {groups: [float(x) for x in f[6].split(',')]}
        else:
            batch_fit.FIT_SPLIT = {groups: [1.0] * groups}
        batch_fit.SPECULATE_BELOW = spec
        for rep in range(3 if first else 2):
            n0 = eng.n_eval
            timer = batch_fit.KernelTimer()
            torch.cuda.synchronize()
            t0 = time.time()
            res = batch_fit.process_batch(None, starts, config=cfg, options={'npoly': 10},
                                          engine=eng, groups=groups, timer=timer, threads=threads, peel=peel)
            torch.cuda.synchronize()
            dt = time.time() - t0
            ks = timer.summary()
            iv = timer.intervals('fused_eval')
            ev = [(r[2], r[1] - r[0]) for r in iv]
            hist = []
            for lo, hi in ((0, 64), (64, 256), (256, 1024), (1024, 4096), (4096, 1 << 30)):
                sel = [d for k, d in ev if lo <= k < hi]
                hist.append(f'[{lo},{hi if hi < 1 << 30 else "inf"}): n={len(sel)} sum={sum(sel):.0f}ms mean={np.mean(sel) if sel else 0:.3f}')
            print('   calls by items:', '; '.join(hist), flush=True)
            if os.environ.get('RVS_TIMELINE') and rep >= 1:
                names = sorted(set(r[0] for r in timer.rec) | set(timer.native))
                ivs = [timer.intervals(n) for n in names]
                allv = np.concatenate(ivs)
                np.savez(os.environ['RVS_TIMELINE'], names=np.array(names),
                         kind=np.concatenate([np.full(len(v), i) for i, v in enumerate(ivs)]),
                         t0=allv[:, 0] - allv[:, 0].min(), t1=allv[:, 1] - allv[:, 0].min(),
                         items=allv[:, 2].astype(np.int64), wall=dt)
            chk = (float(np.sum([r['vel'] for r in res])), float(np.sum([r['chisq'] for r in res])))
            print(f'   spawned {batch_fit.process_batch.last_spawned}  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB')
            print(f'{s} rep{rep}: chk {chk[0]:.9f} {chk[1]:.6f} {B / dt:8.1f} fits/s  {dt:6.3f} s  evals/fit {(eng.n_eval - n0) / B:7.1f}  '
                  f'calls {ks.get("fused_eval_launches")} items/call {ks.get("fused_eval_items_per_launch", 0):.0f} '
                  f'busy {ks.get("fused_eval_ms_busy", 0):.0f} ms  sum {ks.get("fused_eval_ms_total", 0):.0f} ms  '
                  f'scan {ks.get("scan_ms_total", 0):.0f} ms build {ks.get("build_ms_total", 0):.0f} ms  '
                  f'phases { {k: round(v, 2) for k, v in batch_fit.process_batch.last_phase_seconds.items()} }',
                  flush=True)
        first = False
    v = np.array([r['vel'] for r in res])
    print('vel rms vs truth', float(np.sqrt(np.mean((v - vel)**2))), 'median |dv|',
          float(np.median(np.abs(v - vel))))
    if os.environ.get('RVS_PROFILE'):
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        batch_fit.process_batch(None, starts, config=cfg, options={'npoly': 10}, engine=eng,
                                groups=groups, threads=False)
        pr.disable()
        pstats.Stats(pr).sort_stats('tottime').print_stats(35)


if __name__ == '__main__':
    main()
