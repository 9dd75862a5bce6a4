"""Are evaluation calls bit-reproducible under concurrency?  Several threads submit the
same requests over and over (sizes that take the graph route, the direct route and the
chunked route) while another thread runs RV scans; every answer must equal the first."""
import os
import sys
import threading
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from rvspecfit_b200 import spec_fit, spec_inter, batch_fit

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 150


def main():
    import torch
    cfg = bench.make_config(bench.WORKLOADS['desi'])
    setups, objects, pars, vel = bench.make_inputs('desi', B, 1000)
    for st in setups:
        spec_inter.register_bank(spec_inter.bank_from_setup(st), template_lib='synthetic/')
    objs = [[spec_fit.SpecData(*a) for a in o] for o in objects]
    eng = spec_fit.LikelihoodEngine(objs, cfg, {'npoly': 10})
    starts = [dict(bench.FIT_START) for _ in range(B)]
    names = list(spec_inter.getSpecParams(setups[0]['name'], cfg))
    fobj = batch_fit.BatchObjective(eng, names, starts, [], True, cfg, None)
    rs = np.random.RandomState(5)
    reqs = []
    for K in (37, 300, 1500, 5000, 14336, 20000):
        idx = rs.randint(0, B, K)
        X = np.column_stack([vel[idx] + rs.normal(0, 30, K), np.abs(rs.normal(10, 8, K)),
                             rs.uniform(4200, 7000, K), rs.uniform(0.5, 5, K),
                             rs.uniform(-2.3, 0.3, K), rs.uniform(-0.1, 0.6, K)])
        reqs.append((idx, X))
    ref = [fobj(i, X) for i, X in reqs]
    ref2 = [fobj(i, X) for i, X in reqs]
    print('serial repeat identical:', [bool(np.array_equal(a, b)) for a, b in zip(ref, ref2)])
    bad = []
    stop = [False]

    def worker(t):
        order = np.random.RandomState(t).permutation(len(reqs) * REPS) % len(reqs)
        for n, j in enumerate(order):
            got = fobj(*reqs[j])
            if not np.array_equal(got, ref[j]):
                d = np.nonzero(got != ref[j])[0]
                bad.append((t, n, j, len(d), float(np.max(np.abs(got[d] - ref[j][d]) / np.abs(ref[j][d])))))

    vg = np.arange(-1500, 1500, 5.0)
    rag = [np.arange(-40 - i % 7, 40 + i % 5, 0.5 + 0.01 * (i % 3)) + vel[i] for i in range(256)]
    nvr = np.array([len(g) for g in rag])
    Vr = np.zeros((256, nvr.max()))
    for i, g in enumerate(rag):
        Vr[i, :len(g)] = g
        Vr[i, len(g):] = g[-1]
    scan_reqs = [(np.arange(256), np.tile(vg, (256, 1)), np.full(256, len(vg))),
                 (np.arange(256), Vr, nvr)]
    scan_ref = [eng.scan(o, V, nv, pars[:256], fobj.vsini0[:256])[0] for o, V, nv in scan_reqs]
    nscan = [0]

    def scanner(t):
        while not stop[0]:
            for j, (o, V, nv) in enumerate(scan_reqs):
                got = eng.scan(o, V, nv, pars[:256], fobj.vsini0[:256])[0]
                nscan[0] += 1
                if not np.array_equal(got, scan_ref[j], equal_nan=True):
                    d = np.nonzero(~((got == scan_ref[j]) | (np.isnan(got) & np.isnan(scan_ref[j]))))
                    bad.append((100 + t, nscan[0], j, len(d[0]), float(np.nanmax(np.abs(got - scan_ref[j])))))
    ths = [threading.Thread(target=worker, args=(t,)) for t in range(3)]
    sc = threading.Thread(target=scanner, args=(0,))
    sc2 = threading.Thread(target=scanner, args=(1,))
    t0 = time.time()
    sc.start()
    sc2.start()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    stop[0] = True
    sc.join()
    sc2.join()
    print('scans compared:', nscan[0])
    print(f'{3 * len(reqs) * REPS} concurrent calls in {time.time() - t0:.1f} s; mismatching calls: {len(bad)}')
    for b in bad[:20]:
        print('  thread %d call %d request %d: %d items differ, max rel %.3g' % b)


if __name__ == '__main__':
    main()
