"""Batched drivers over spec_fit.LikelihoodEngine: many objects stepped together
so that every likelihood evaluation of a step is one kernel launch per arm.
"""
import numpy as np

from . import _dev, spec_fit


class KernelTimer:
    """CUDA-event timing of individual launches on the launching stream.
    Attach with engine.timer = KernelTimer(); read summary() after the run."""

    def __init__(self):
        self.rec = []

    def reset(self):
        self.rec = []

    def start(self):
        torch = _dev.torch_mod()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return e0

    def stop(self, name, e0, items):
        torch = _dev.torch_mod()
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.rec.append((name, e0, e1, items))

    def summary(self):
        torch = _dev.torch_mod()
        torch.cuda.synchronize()
        out = {}
        for name in sorted(set(r[0] for r in self.rec)):
            rs = [r for r in self.rec if r[0] == name]
            ms = [r[1].elapsed_time(r[2]) for r in rs]
            out[name + '_launches'] = len(rs)
            out[name + '_ms_total'] = float(np.sum(ms))
            out[name + '_ms_per_launch'] = float(np.mean(ms))
            out[name + '_items_per_launch'] = float(np.mean([r[3] for r in rs]))
        return out


def scan_and_evaluate(eng, start, vgrid, tp, tv, tvs, timer=None, groups=2):
    """One pass of the hot path over the engine's objects: an RV-grid scan at
    the start parameters with find_best statistics, then len(tp) rounds of
    optimiser-phase evaluations, each at new template parameters.  The objects
    are stepped as `groups` independent groups in ping-pong (LikelihoodEngine.
    submit / result): while the host digests the results of one group and
    prepares its next trial points, the GPU works on the other.  Returns (B, 6):
    best_chi, best_vel, vel_err, skewness, kurtosis of the scan and the smallest
    chi-square met in the evaluation rounds."""
    B = eng.nobj
    obj = np.arange(B)
    eng.timer = timer
    chi = eng.evaluate(obj, np.tile(vgrid, (B, 1)), start, None)
    st, _ = spec_fit.scan_stats(np.tile(vgrid, (B, 1)), chi[:, None, :])
    best = np.full(B, np.inf)
    parts = [p for p in np.array_split(obj, max(1, min(groups, B))) if len(p)]
    pending = []
    for e in range(len(tp)):
        for p in parts:
            if len(pending) >= len(parts):
                q, h = pending.pop(0)
                best[q] = np.minimum(best[q], h.result())
            pending.append((p, eng.submit(p, tv[e][p], tp[e][p], tvs[e][p])))
    for q, h in pending:
        best[q] = np.minimum(best[q], h.result())
    eng.timer = None
    return np.concatenate([st[:, :5], best[:, None]], axis=1)
