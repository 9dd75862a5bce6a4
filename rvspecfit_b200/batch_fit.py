"""Batched drivers over spec_fit.LikelihoodEngine: many objects stepped together
so that every likelihood evaluation of a step is one kernel launch per arm.

`process_batch` is `vel_fit.process` (reference vel_fit.py:505-737) for a list of
objects.  Every object follows exactly the decision
rules of the single-object path:
  * the RV-grid scan and its refinement (`_minimum_sampler`, vel_fit.py:358-439)
    run with all active objects' grids in one launch per round;
  * Nelder-Mead is scipy's algorithm (scipy/optimize/_optimize.py,
    `_minimize_neldermead`: rho, chi, psi, sigma = 1, 2, 1/2, 1/2, the same
    termination test, the same acceptance rules, the same ordering), written for
    a whole batch of simplices in lock-step -- each iteration needs the
    reflection point of every active object (one launch), then the expansion /
    contraction points of those that ask for one (one launch), then the shrunk
    simplices (one launch);
  * BFGS is scipy's `_minimize_bfgs` with its MINPACK line search, restated
    over arrays of problems (batch_bfgs.bfgs_steps): one function value and one
    forward-difference gradient per live problem and round, in one launch;
  * the Hessian is the same central-difference stencil as vel_fit's, all
    objects' points in one launch.
The objects are split into a few lock-step sets, each a coroutine on a host thread of
its own (fit_steps, run_threads).  The rounds of both optimiser stages are stepped by
the library without the interpreter (rvs_nm_drive / rvs_bfgs_drive on an evaluation slot
the stage holds: stepper, packing, captured graph launch, wait, reduction), and the
objects that have finished a stage move on in groups while the slower ones iterate, so
the host logic of one set and the latency-bound tail of a stage (few live problems) run
beside the large calls of the rest.  threads=False keeps everything on the calling thread,
rounds driven from Python (NMStepper / batch_bfgs.bfgs_steps, which visit scipy's points
bit for bit).
"""
import os

import numpy as np

from . import _dev, batch_bfgs, spec_fit, spec_inter, vel_fit


class KernelTimer:
    """CUDA-event timing of individual launches on the launching stream.
    Attach with engine.timer = KernelTimer(); read summary() after the run.
    Launches made by the native round loop (rvs_nm_drive) arrive as arrays of
    (start, end, items) in ms since the timer's epoch event (add_native)."""

    def __init__(self):
        torch = _dev.torch_mod()
        self.rec = []
        self.native = {}
        self.epoch = torch.cuda.Event(enable_timing=True)
        self.epoch.record()
        self.epoch.synchronize()

    def reset(self):
        """Drop the records and restart the clock (event times are single-precision ms
        since the epoch: a fresh epoch keeps them fine-grained in long runs)."""
        self.rec = []
        self.native = {}
        _dev.torch_mod().cuda.synchronize()
        self.epoch.record()
        self.epoch.synchronize()

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

    def add_native(self, name, rows):
        self.native.setdefault(name, []).append(rows)

    def intervals(self, name):
        """(n, 3) array: start, end (ms since the epoch), items of every launch `name`."""
        _dev.torch_mod().cuda.synchronize()
        rows = [(self.epoch.elapsed_time(r[1]), self.epoch.elapsed_time(r[2]), r[3])
                for r in list(self.rec) if r[0] == name]
        parts = [np.array(rows, dtype=np.float64).reshape(-1, 3)] + list(self.native.get(name, []))
        return np.concatenate(parts)

    def summary(self):
        out = {}
        for name in sorted(set(r[0] for r in list(self.rec)) | set(self.native)):
            iv = self.intervals(name)
            ms = iv[:, 1] - iv[:, 0]
            out[name + '_launches'] = len(iv)
            out[name + '_ms_total'] = float(np.sum(ms))
            out[name + '_ms_per_launch'] = float(np.mean(ms))
            out[name + '_items_per_launch'] = float(np.mean(iv[:, 2]))
            # launches of concurrent lock-step sets overlap on the device: the time
            # during which at least one of them was running
            busy, end = 0.0, -np.inf
            for a, b, _ in iv[np.argsort(iv[:, 0], kind='stable')]:
                if b > end:
                    busy += b - max(a, end)
                    end = b
            out[name + '_ms_busy'] = float(busy)
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
    # device time of the whole evaluation phase: the calls of different groups
    # overlap on the GPU (each in-flight evaluation has its own streams), so
    # their individual durations do not add up
    e0 = timer.start() if timer is not None else None
    for e in range(len(tp)):
        for p in parts:
            if len(pending) >= len(parts):
                q, h = pending.pop(0)
                best[q] = np.minimum(best[q], h.result())
            pending.append((p, eng.submit(p, tv[e][p], tp[e][p], tvs[e][p])))
    for q, h in pending:
        best[q] = np.minimum(best[q], h.result())
    if timer is not None:
        timer.stop('eval_phase', e0, len(tp) * B)
    eng.timer = None
    return np.concatenate([st[:, :5], best[:, None]], axis=1)


# ------------------------------------------------------------ batched objective
class BatchObjective:
    """chisq_func / chisq_func0 of vel_fit (reference vel_fit.py:210-257) for
    many (object, parameter-vector) pairs at once.  All objects share the layout
    of the fitted vector: [vel, (vsini), free atmospheric parameters]."""

    def __init__(self, eng, specParams, paramDict0s, fixParam, fitVsini, config, priors=None):
        self.eng = eng
        self.specParams = list(specParams)
        self.fix = [p in fixParam for p in self.specParams]
        self.fitVsini = fitVsini
        self.p0 = np.array([[d[p] for p in self.specParams] for d in paramDict0s], dtype=np.float64)
        self.has_vsini = 'vsini' in paramDict0s[0]
        self.vsini0 = np.array([d.get('vsini', 0.0) for d in paramDict0s], dtype=np.float64)
        self.vsini_fixed = self.has_vsini and not fitVsini
        self.min_vel, self.max_vel = config['min_vel'], config['max_vel']
        self.max_vsini = config['max_vsini']
        self.priors = priors
        self.nfev = 0
        self._lay = None

    def layout(self):
        """struct rvs_fit_layout of this objective (the arrays it points to are kept
        alive here), or False when the engine has no packed fast path."""
        if self._lay is None:
            eng = self.eng
            self._lay = False
            if hasattr(eng, 'submit_fit') and len(self.specParams) <= 30:
                from . import _cabi
                bank0 = eng.arms[eng.setups[0]]['bank']
                ns = len(self.specParams)
                lay = _cabi.FitLayout()
                lay.nfit = 1 + int(self.fitVsini) + sum(not f for f in self.fix)
                lay.nspec, lay.fit_vsini = ns, int(self.fitVsini)
                lay.has_vsini = int(self.has_vsini)
                lay.fixmask = sum(1 << j for j, f in enumerate(self.fix) if f)
                lay.logmask = sum(1 << j for j in bank0.log_ids)
                pri = self.priors or {}
                lay.priormask = sum(1 << j for j, k in enumerate(self.specParams) if k in pri)
                lay.narm, lay.nobj = len(eng.setups), eng.nobj
                lay.min_vel, lay.max_vel = float(self.min_vel), float(self.max_vel)
                lay.max_vsini = float(self.max_vsini)
                keep = dict(
                    p0=np.ascontiguousarray(self.p0),
                    q0=np.ascontiguousarray(spec_inter.map_params(self.p0, bank0.log_ids)),
                    vs0=np.ascontiguousarray(self.vsini0),
                    mu=np.array([pri[k][0] if k in pri else 0.0 for k in self.specParams],
                                dtype=np.float64),
                    sig=np.array([pri[k][1] if k in pri else 1.0 for k in self.specParams],
                                 dtype=np.float64),
                    oix=np.ascontiguousarray(eng._oix, dtype=np.int32),
                    badchi=np.ascontiguousarray(eng.badchi, dtype=np.float64),
                    cover=np.ascontiguousarray(eng._cover0, dtype=np.uint8))
                lay.h_p0, lay.h_q0 = keep['p0'].ctypes.data, keep['q0'].ctypes.data
                lay.h_vsini0 = keep['vs0'].ctypes.data
                lay.h_prior_mu, lay.h_prior_sig = keep['mu'].ctypes.data, keep['sig'].ctypes.data
                lay.h_oix, lay.h_badchi = keep['oix'].ctypes.data, keep['badchi'].ctypes.data
                lay.h_cover = keep['cover'].ctypes.data
                # fitted columns of the log-mapped parameters, in parameter order
                cols, pos = [], 1 + int(self.fitVsini)
                for j, fixed in enumerate(self.fix):
                    if not fixed:
                        if j in bank0.log_ids:
                            cols.append(pos)
                        pos += 1
                self._lay, self._lay_keep, self._logcols = lay, keep, cols
        return self._lay

    def unpack(self, idx, X):
        """ParamMapper.forward for a batch: vel (K,), vsini (K,) or None,
        params (K, nspec), penalty (K,)."""
        X = np.asarray(X, dtype=np.float64)
        vel = X[:, 0]
        pos = 1
        pen = np.zeros(len(X))
        vsini = None
        if self.fitVsini:
            xv = X[:, 1]
            vsini = np.clip(xv, 0, self.max_vsini)
            pen = (xv < 0) * (vsini - xv)**2 + (xv > self.max_vsini) * (vsini - xv)**2
            pos = 2
        elif self.vsini_fixed:
            vsini = self.vsini0[idx]
        params = np.empty((len(X), len(self.specParams)))
        for j, fixed in enumerate(self.fix):
            if fixed:
                params[:, j] = self.p0[idx, j]
            else:
                params[:, j] = X[:, pos]
                pos += 1
        assert pos == X.shape[1]
        return vel, vsini, params, pen

    def prior_term(self, params):
        tot = np.zeros(len(params))
        if self.priors is not None:
            for i, k in enumerate(self.specParams):
                if k in self.priors:
                    tot = tot + ((self.priors[k][0] - params[:, i]) / self.priors[k][1])**2
        return tot

    def chisq0(self, idx, vel, vsini, params):
        """chisq_func0: priors + -2 log L."""
        return self.submit0(idx, vel, vsini, params)()

    def submit0(self, idx, vel, vsini, params):
        """Start chisq_func0 (vel_fit.py:210-230) for K (object, point) pairs; returns
        a waiter."""
        if len(idx) > MAX_CALL_ITEMS:
            return _ChunkedWaiter(
                len(idx), lambda a, b: self.submit0(idx[a:b], vel[a:b],
                                                    None if vsini is None else vsini[a:b],
                                                    params[a:b]))
        idx = np.asarray(idx, dtype=np.int64)
        self.nfev += len(idx)
        return _Waiter(self._start(idx, vel, params, vsini), None, self.prior_term(params), 0.0,
                       len(idx))

    def __call__(self, idx, X):
        """chisq_func for K pairs: idx (K,) object indices, X (K, N)."""
        return self.submit(idx, X)()

    def _start(self, idx, vel, params, vsini):
        if hasattr(self.eng, 'submit'):
            return self.eng.submit(idx, vel, params, vsini)
        return _Done(self.eng.evaluate(idx, vel, params, vsini))    # blocking engines

    def redo_values(self, obj32, X):
        """chisq_func through the engine's general path (items the fused path could not
        settle: off-grid templates that are not usable, a normal matrix that is not
        positive definite, ...)."""
        idx = np.asarray(obj32, dtype=np.int64)
        vel, vsini, params, pen = self.unpack(idx, X)
        eng = self.eng
        with eng._general_lock:
            chi = eng._evaluate_general(idx, vel, params, vsini, True, None, False, False)
        return self.prior_term(params) + chi + pen

    def submit(self, idx, X):
        """Start the evaluation of chisq_func for K pairs and return a waiter: a
        callable that gives the values (priors + -2 log L + penalty, 1e30 behind the
        hard walls, vel_fit.py:210-257), with ready() telling whether it would block."""
        if len(idx) > MAX_CALL_ITEMS:
            return _ChunkedWaiter(len(idx), lambda a, b: self.submit(idx[a:b], X[a:b]))
        lay = self.layout()
        if lay:
            obj32 = np.ascontiguousarray(idx, dtype=np.int32)
            X = np.ascontiguousarray(X, dtype=np.float64)
            with np.errstate(all='ignore'):
                logvals = np.log10(X[:, self._logcols].T) if self._logcols else None
            pend = self.eng.submit_fit(lay, obj32, X,
                                       None if logvals is None else np.ascontiguousarray(logvals))
            if pend is not None:
                self.nfev += len(obj32)
                return _FitWaiter(self, pend, obj32, X)
        idx = np.asarray(idx, dtype=np.int64)
        vel, vsini, params, pen = self.unpack(idx, X)
        wall = (vel > self.max_vel) | (vel < self.min_vel) | ~np.isfinite(params).all(axis=1)
        ok = ~wall
        if ok.all():
            self.nfev += len(idx)
            return _Waiter(self._start(idx, vel, params, vsini), None, self.prior_term(params),
                           pen, len(idx))
        pend = None
        if ok.any():
            self.nfev += int(ok.sum())
            pend = self._start(idx[ok], vel[ok], params[ok], None if vsini is None else vsini[ok])
        return _Waiter(pend, ok, self.prior_term(params[ok]), pen[ok], len(idx))


# Largest evaluation call: longer requests (the Hessian stencils of a large group) go out
# in pieces, two in flight, so that the per-call device buffers stay bounded
MAX_CALL_ITEMS = 16384


class _ChunkedWaiter:
    """Values of a request served by several evaluation calls, two of them in flight."""

    def __init__(self, n, submit):
        self.n, self.sub = n, submit
        self.cuts = list(range(0, n, MAX_CALL_ITEMS)) + [n]
        self.pend = [submit(self.cuts[i], self.cuts[i + 1])
                     for i in range(min(2, len(self.cuts) - 1))]

    def ready(self):
        return False

    def __call__(self):
        out = np.empty(self.n)
        nxt = len(self.pend)
        for i in range(len(self.cuts) - 1):
            out[self.cuts[i]:self.cuts[i + 1]] = self.pend.pop(0)()
            if nxt < len(self.cuts) - 1:
                self.pend.append(self.sub(self.cuts[nxt], self.cuts[nxt + 1]))
                nxt += 1
        return out


class _Done:
    def __init__(self, v):
        self.v = v

    def result(self):
        return self.v

    def ready(self):
        return True


class _FitWaiter:
    """Values of an evaluation started through LikelihoodEngine.submit_fit; the items
    the fused path could not settle go through the general path here."""

    def __init__(self, fobj, pend, obj32, X):
        self.fobj, self.pend, self.obj32, self.X = fobj, pend, obj32, X

    def ready(self):
        return self.pend.ready()

    def __call__(self):
        out, redo = self.pend.result()
        if redo is not None:
            r = np.nonzero(redo)[0]
            fobj = self.fobj
            idx = self.obj32[r].astype(np.int64)
            vel, vsini, params, pen = fobj.unpack(idx, self.X[r])
            eng = fobj.eng
            eng.n_eval -= len(r)
            with eng._general_lock:
                chi = eng._evaluate_general(idx, vel, params, vsini, True, None, False, False)
            out[r] = fobj.prior_term(params) + chi + pen
        return out


class _Waiter:
    """Values of a started objective evaluation: 1e30 behind the walls (`ok` False),
    additive host terms + the engine's -2 log L elsewhere."""

    def __init__(self, pend, ok, prior, pen, n):
        self.pend, self.ok, self.prior, self.pen, self.n = pend, ok, prior, pen, n

    def ready(self):
        return self.pend is None or self.pend.ready()

    def __call__(self):
        if self.ok is None:
            return self.prior + self.pend.result() + self.pen
        out = np.full(self.n, 1e30)
        if self.pend is not None:
            out[self.ok] = self.prior + self.pend.result() + self.pen
        return out


# below this many active problems an iteration's candidate points go out in one call
SPECULATE_BELOW = 128
# smallest lock-step set worth its own evaluation calls
NM_MIN_GROUP = 64


# ------------------------------------------------------- lock-step Nelder-Mead
def nelder_mead_steps(sims, xatol=1e-2, fatol=1e-3, maxiter=10000, speculate_below=0):
    """scipy's Nelder-Mead (`_minimize_neldermead`, adaptive=False, no bounds,
    maxfev=inf) on B simplices at once, as a generator: it yields evaluation
    requests (idx (K,), X (K, N)) -- rows X belonging to problems idx -- is sent
    their function values, and returns dict(x (B, N), fun (B,), success (B,),
    final_simplex (B, N+1, N), nit (B,), nfev (B,)).  sims (B, N+1, N): initial
    simplices.  Problem b visits exactly the points, in exactly the order, that
    scipy.optimize.minimize(method='Nelder-Mead', options={'initial_simplex':
    sims[b], ...}) would.  Drivers: nelder_mead_lockstep (one generator, blocking
    evaluations), nelder_mead_interleaved (several generators whose host work
    overlaps each other's evaluations on the GPU)."""
    sim = np.array(sims, dtype=np.float64)
    B, N1, N = sim.shape
    assert N1 == N + 1
    rho, chi, psi, sigma = 1, 2, 0.5, 0.5
    allb = np.arange(B)
    fsim = (yield np.repeat(allb, N1), sim.reshape(B * N1, N)).reshape(B, N1)
    nfev = np.full(B, N1)

    def sort(rows):
        ind = np.argsort(fsim[rows], axis=1, kind='stable')
        fsim[rows] = np.take_along_axis(fsim[rows], ind, axis=1)
        sim[rows] = np.take_along_axis(sim[rows], ind[:, :, None], axis=1)
    sort(allb)
    iterations = np.ones(B, dtype=np.int64)
    active = np.ones(B, dtype=bool)
    success = np.zeros(B, dtype=bool)
    while active.any():
        a = np.nonzero(active)[0]
        # termination tests of the while-loop head and of its first statement
        out_of_iters = iterations[a] >= maxiter
        conv = (np.max(np.abs(sim[a, 1:] - sim[a, :1]).reshape(len(a), -1), axis=1) <= xatol) & \
            (np.max(np.abs(fsim[a, :1] - fsim[a, 1:]), axis=1) <= fatol)
        stop = out_of_iters | conv
        success[a[conv & ~out_of_iters]] = True
        active[a[stop]] = False
        a = a[~stop]
        if len(a) == 0:
            break
        xbar = np.add.reduce(sim[a, :-1], 1) / N
        last = sim[a, -1]
        xr = (1 + rho) * xbar - rho * last
        f0, fm2, fm1 = fsim[a, 0], fsim[a, -2], fsim[a, -1]
        if len(a) <= speculate_below:
            # few problems left: launches are latency-bound, so the three candidate
            # second points are evaluated together with the reflection in ONE call;
            # each problem then uses exactly the value scipy would have computed
            xe = (1 + rho * chi) * xbar - rho * chi * last
            xc = (1 + psi * rho) * xbar - psi * rho * last
            xcc = (1 - psi) * xbar + psi * last
            fall = (yield np.tile(a, 4), np.concatenate([xr, xe, xc, xcc])).reshape(4, len(a))
            fxr = fall[0]
            nfev[a] += 1
            expand = fxr < f0
            accept_r = ~expand & (fxr < fm2)
            contract = ~expand & ~accept_r & (fxr < fm1)
            inside = ~expand & ~accept_r & ~contract
            x2 = np.where(expand[:, None], xe, np.where(contract[:, None], xc, xcc))
            f2 = np.where(expand, fall[1], np.where(contract, fall[2], fall[3]))
            need2 = expand | contract | inside
            f2[~need2] = np.nan
            nfev[a[need2]] += 1
        else:
            fxr = yield a, xr
            nfev[a] += 1
            expand = fxr < f0
            accept_r = ~expand & (fxr < fm2)
            contract = ~expand & ~accept_r & (fxr < fm1)
            inside = ~expand & ~accept_r & ~contract
            # second point, where one is asked for
            x2 = np.empty_like(xr)
            x2[expand] = (1 + rho * chi) * xbar[expand] - rho * chi * last[expand]
            x2[contract] = (1 + psi * rho) * xbar[contract] - psi * rho * last[contract]
            x2[inside] = (1 - psi) * xbar[inside] + psi * last[inside]
            need2 = expand | contract | inside
            f2 = np.full(len(a), np.nan)
            if need2.any():
                f2[need2] = yield a[need2], x2[need2]
                nfev[a[need2]] += 1
        new_x, new_f = xr.copy(), fxr.copy()
        take2 = (expand & (f2 < fxr)) | (contract & (f2 <= fxr)) | (inside & (f2 < fm1))
        new_x[take2], new_f[take2] = x2[take2], f2[take2]
        shrink = (contract & ~(f2 <= fxr)) | (inside & ~(f2 < fm1))
        keep = ~shrink
        sim[a[keep], -1] = new_x[keep]
        fsim[a[keep], -1] = new_f[keep]
        if shrink.any():
            s = a[shrink]
            sim[s, 1:] = sim[s, :1] + sigma * (sim[s, 1:] - sim[s, :1])
            fsim[s, 1:] = (yield np.repeat(s, N), sim[s, 1:].reshape(len(s) * N, N)).reshape(len(s), N)
            nfev[s] += N
        iterations[a] += 1
        sort(a)
    return dict(x=sim[:, 0].copy(), fun=np.min(fsim, axis=1), success=success,
                final_simplex=sim, nit=iterations, nfev=nfev)


class NMStepper:
    """The library's host-side lock-step Nelder-Mead stepper (csrc/nm_host.cpp, rvs_nm_*)
    for B simplices: request() -> (idx, X) or None, feed(values), active(), result().
    A stopped problem's rows of result() are final while the others go on."""

    def __init__(self, sims, xatol=1e-2, fatol=1e-3, maxiter=10000):
        import ctypes
        from . import _cabi
        self.L = _cabi.lib()
        self.sim = np.array(sims, dtype=np.float64, order='C')   # receives the final simplices
        self.B, N1, self.N = self.sim.shape
        assert N1 == self.N + 1
        h = self.L.rvs_nm_create(self.B, self.N, _dev.hptr(self.sim), float(xatol), float(fatol),
                                 int(maxiter))
        if not h:
            raise _cabi.RvsError('rvs_nm_create failed')
        self.h = ctypes.c_void_p(h)
        self.cap = self.B * max(N1, 4)
        self.idx = np.empty(self.cap, dtype=np.int32)
        self.X = np.empty((self.cap, self.N), dtype=np.float64)
        self.n = 0

    def request(self, speculate_below=0):
        n = self.L.rvs_nm_request(self.h, int(speculate_below), _dev.hptr(self.idx),
                                  _dev.hptr(self.X), self.cap)
        assert n <= self.cap
        self.n = n
        return (self.idx[:n], self.X[:n]) if n else None

    def feed(self, f):
        from . import _cabi
        f = np.ascontiguousarray(f, dtype=np.float64)
        _cabi.check(self.L.rvs_nm_feed(self.h, _dev.hptr(f), self.n), 'rvs_nm_feed')

    def active(self):
        a = np.empty(self.B, dtype=np.uint8)
        self.L.rvs_nm_live(self.h, _dev.hptr(a))
        return a.astype(bool)

    def result(self):
        from . import _cabi
        B, N = self.B, self.N
        x, fun = np.empty((B, N)), np.empty(B)
        success = np.empty(B, dtype=np.uint8)
        nit, nfev = np.empty(B, dtype=np.int64), np.empty(B, dtype=np.int64)
        sim = np.empty_like(self.sim)
        _cabi.check(self.L.rvs_nm_result(self.h, _dev.hptr(x), _dev.hptr(fun), _dev.hptr(success),
                                         _dev.hptr(sim), _dev.hptr(nit), _dev.hptr(nfev)),
                    'rvs_nm_result')
        return dict(x=x, fun=fun, success=success.astype(bool), final_simplex=sim, nit=nit,
                    nfev=nfev)

    def close(self):
        if self.h is not None:
            self.L.rvs_nm_destroy(self.h)
            self.h = None


def nelder_mead_native(sims, xatol=1e-2, fatol=1e-3, maxiter=10000, speculate_below=0):
    """nelder_mead_steps with the stepping done by the library's host-side stepper
    (NMStepper): the same generator protocol, the same trajectories
    (tests/test_batch_drivers.py compares both with scipy), a fraction of the host
    time per round.  `speculate_below` may be a callable giving the threshold for the
    coming round."""
    st = NMStepper(sims, xatol, fatol, maxiter)
    try:
        while True:
            req = st.request(speculate_below() if callable(speculate_below) else speculate_below)
            if req is None:
                break
            st.feed((yield req))
        return st.result()
    finally:
        st.close()


def nelder_mead_lockstep(fbatch, sims, xatol=1e-2, fatol=1e-3, maxiter=10000,
                         speculate_below=0, native=False):
    """nelder_mead_steps (or its native sibling) driven with a blocking objective
    fbatch(idx, X) -> f."""
    gen = (nelder_mead_native if native else nelder_mead_steps)(sims, xatol, fatol, maxiter,
                                                                speculate_below)
    try:
        req = next(gen)
        while True:
            req = gen.send(fbatch(*req))
    except StopIteration as stop:
        return stop.value


def nelder_mead_interleaved(fsubmit, sims, groups=2, xatol=1e-2, fatol=1e-3, maxiter=10000,
                            speculate_below=0):
    """The same minimisations with the problems split into `groups` independent
    lock-step sets.  fsubmit(idx, X) starts an evaluation and returns a callable
    that waits for its values; while one set's request is on the GPU the host
    advances the others, so device and host work overlap.  Every problem follows
    its own trajectory, so the result equals nelder_mead_lockstep's."""
    sims = np.asarray(sims, dtype=np.float64)
    B = len(sims)
    parts = [p for p in np.array_split(np.arange(B), max(1, min(groups, B))) if len(p)]
    gens = [nelder_mead_steps(sims[p], xatol, fatol, maxiter, speculate_below) for p in parts]
    out = [None] * len(parts)
    pending = []
    for gi, gen in enumerate(gens):
        idx, X = next(gen)      # a generator always asks for its initial simplices
        pending.append((gi, fsubmit(parts[gi][idx], X)))
    while pending:
        gi, wait = pending.pop(0)
        try:
            idx, X = gens[gi].send(wait())
            pending.append((gi, fsubmit(parts[gi][idx], X)))
        except StopIteration as stop:
            out[gi] = stop.value
    res = {}
    for k in out[0]:
        full = np.empty((B,) + out[0][k].shape[1:], dtype=out[0][k].dtype)
        for p, o in zip(parts, out):
            full[p] = o[k]
        res[k] = full
    return res


# ------------------------------------------------------------ pipeline driver
class _Ready:
    """Waiter of a request that was served when it was made."""

    def __init__(self, value):
        self.value = value

    def ready(self):
        return True

    def __call__(self):
        return self.value


def run_pipeline(gens, start):
    """Advance several request-yielding coroutines, keeping one request of each
    in flight.  `start(request)` begins serving a request and returns a waiter
    (callable giving the answer, with ready() telling whether it would block).
    Whichever coroutine's answer is ready first is advanced first, so host work
    of one lock-step set overlaps the device work of the others.  A request
    ('spawn', coroutine) adds a coroutine to the set.  Returns the coroutines'
    return values, the given ones first and in order."""
    gens = list(gens)
    out = [None] * len(gens)
    pending = []

    def advance(gi, value, first=False):
        try:
            while True:
                req = next(gens[gi]) if first else gens[gi].send(value)
                first = False
                if req[0] != 'spawn':
                    pending.append((gi, start(req)))
                    return
                gens.append(req[1])
                out.append(None)
                advance(len(gens) - 1, None, True)
                value = None
        except StopIteration as stop:
            out[gi] = stop.value
    for gi in range(len(gens)):
        advance(gi, None, True)
    while pending:
        pick = 0
        for j, (_, w) in enumerate(pending):
            if w.ready():
                pick = j
                break
        gi, wait = pending.pop(pick)
        advance(gi, wait())
    return out


def run_threads(gens, start, device=None, post=None):
    """run_pipeline with one host thread per coroutine.  The per-round host work of a
    lock-step set is library calls that release the interpreter lock (the optimiser
    stepper rvs_nm_*, rvs_fit_pack / rvs_fit_collect, graph launches, event waits), so
    the sets' host work runs on different cores instead of queueing behind each other.
    A request ('spawn', coroutine) starts a thread for that coroutine.
    `device`: CUDA device index the threads make current (a new thread starts on
    device 0).  `post`: applied to every coroutine's return value on its own thread (the
    result assembly of one set then runs under the device work of the others)."""
    import threading
    gens = list(gens)
    out = [None] * len(gens)
    errors = []
    threads = []
    lock = threading.Lock()

    def launch(gen):
        with lock:
            gi = len(threads)
            if gi >= len(out):
                out.append(None)
            t = threading.Thread(target=work, args=(gi, gen), daemon=True)
            threads.append(t)
        t.start()

    def work(gi, gen):
        try:
            if device is not None:
                _dev.torch_mod().cuda.set_device(device)
            req = next(gen)
            while not errors:
                if req[0] == 'spawn':
                    launch(req[1])
                    req = gen.send(None)
                else:
                    req = gen.send(start(req)())
        except StopIteration as stop:
            try:
                out[gi] = stop.value if post is None else post(stop.value)
            except BaseException as exc:      # noqa: BLE001
                errors.append(exc)
        except BaseException as exc:          # noqa: BLE001  (re-raised by the caller)
            errors.append(exc)
    for gen in gens:
        launch(gen)
    i = 0
    while True:
        with lock:
            if i >= len(threads):
                break
            t = threads[i]
        t.join()
        i += 1
    if errors:
        raise errors[0]
    return out


def _drive(gen, sel):
    """Run an optimiser coroutine (requests (idx, X) in its own problem numbers)
    inside a fit coroutine: requests become ('f', objects, X)."""
    try:
        idx, X = next(gen)
        while True:
            idx, X = gen.send((yield ('f', sel[idx], X)))
    except StopIteration as stop:
        return stop.value


# ------------------------------------------------------------ the batched fit
def _scan_general(eng, idx, V, nv, params, vsini):
    """Scan statistics (n, 8) through the engine's general path: chi-square matrix of every
    arm to the host (penalties, SVD rescue, exceptions), then the statistics kernel."""
    lock = getattr(eng, '_general_lock', None)
    import contextlib
    with (lock if lock is not None else contextlib.nullcontext()):
        chi = eng.evaluate(idx, V, params, vsini)
        st, _ = spec_fit.scan_stats(V, chi[:, None, :],
                                    nv=None if (nv == V.shape[1]).all() else nv,
                                    want_probs=False)
    return st


def _scan_round(eng, idx, grids, params, vsini):
    """find_best (one template per object) for ragged velocity grids: chi-squares
    in one launch per arm, statistics on the device per distinct grid length.
    Returns (n, 5): best_chi, best_vel, vel_err, skewness, kurtosis."""
    nv = np.array([len(g) for g in grids])
    nmax = int(nv.max())
    ragged = not (nv == nmax).all()
    if not ragged:
        V = np.stack(grids)
    else:
        V = np.empty((len(grids), nmax))
        for i, g in enumerate(grids):
            V[i, :nv[i]] = g
            V[i, nv[i]:] = g[-1]
    idx = np.asarray(idx)
    fast = eng.scan(idx, V, nv, params, vsini) \
        if hasattr(eng, 'scan') and not os.environ.get('RVS_NO_FAST_SCAN') else None
    if fast is not None:
        # device-resident route; the objects it could not settle take the general one
        st, redo = fast
        if redo.any():
            r = np.nonzero(redo)[0]
            st[r] = _scan_general(eng, idx[r], V[r], nv[r], params[r],
                                  None if vsini is None else np.asarray(vsini)[r])
    else:
        st = _scan_general(eng, idx, V, nv, params, vsini)
    # the reference's find_best would trip its assertion / propagate the NaN here
    # (spec_fit.py:1014, 1072-1092); vel_fit.process raises, and so does the batch
    bad = st[:, 7] != 0
    if bad.any():
        raise RuntimeError('RV scan without a usable minimum (flat or non-finite chi-square '
                           f'around it) for object(s) {idx[bad].tolist()}')
    return st[:, :5]


def hessian_points(x, hs):
    """Evaluation points of vel_fit.central_hessian for every row of x (B, n):
    (B, npts, n), the centre first, then for the step sets hs and hs/2 the
    +-e_i pairs and the four (+-e_i, +-e_j) corners of every j < i, in the
    order that routine asks for them."""
    x = np.asarray(x, dtype=np.float64)
    B, n = x.shape
    hs = np.asarray(hs, dtype=np.float64)
    pts = [x]
    for h in (hs / 2, hs):
        for i in range(n):
            for si in (1, -1):
                p = x.copy()
                p[:, i] = x[:, i] + si * h[i]
                pts.append(p)
            for j in range(i):
                for si, sj in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
                    p = x.copy()
                    p[:, i] = x[:, i] + si * h[i]
                    p[:, j] = x[:, j] + sj * h[j]
                    pts.append(p)
    return np.stack(pts, axis=1)


def hessian_from_values(vals, hs):
    """vel_fit.central_hessian's arithmetic on the values at hessian_points:
    vals (B, npts) -> (B, n, n)."""
    hs = np.asarray(hs, dtype=np.float64)
    n = len(hs)
    B = len(vals)
    f0 = vals[:, 0]
    pos = 1
    both = []
    for h in (hs / 2, hs):
        H = np.zeros((B, n, n))
        for i in range(n):
            fp, fm = vals[:, pos], vals[:, pos + 1]
            pos += 2
            H[:, i, i] = (fp - 2 * f0 + fm) / h[i]**2
            for j in range(i):
                fpp, fpm, fmp, fmm = (vals[:, pos + k] for k in range(4))
                pos += 4
                H[:, i, j] = H[:, j, i] = (fpp - fpm - fmp + fmm) / (4 * h[i] * h[j])
        both.append(H)
    return (4 * both[0] - both[1]) / 3


def simplex_starts(best_vel, fobj, specParams, fixParam, fitVsini, max_vsini):
    """vel_fit._get_simplex_start for every object: (B, N + 1, N)."""
    cols, std = [best_vel], [5]
    if fitVsini:
        cols.append(np.clip(fobj.vsini0, 0, max_vsini))
        std.append(3)
    for j, name in enumerate(specParams):
        if name not in fixParam:
            cols.append(fobj.p0[:, j])
            std.append(vel_fit.SIMPLEX_STD.get(name) or 0.5)
    cur = np.stack(cols, axis=1)
    std = np.array(std)
    ndim = cur.shape[1]
    R = np.random.RandomState(43434)        # the reference's fixed seed, vel_fit.py:306
    noise = std[None, :] * R.normal(size=(ndim, ndim))
    sims = np.empty((len(cur), ndim + 1, ndim))
    sims[:, 0] = cur
    sims[:, 1:] = cur[:, None, :] + noise[None]
    return sims


# A lock-step stage hands the problems that have finished on to the next stage of the
# fit as soon as that many of them wait (and at least this fraction of the stage's
# problems), instead of keeping them until the slowest one stops
PEEL_MIN = 384
PEEL_FRAC = 0.3


class _FitCtx:
    """What the stages of one lock-step set share."""

    def __init__(self, sel, fobj, specParams, fixParam, fitVsini, has_vsini, config, phase, peel,
                 threaded=False):
        self.sel, self.fobj, self.specParams, self.fixParam = sel, fobj, specParams, fixParam
        self.fitVsini, self.has_vsini, self.config, self.phase = fitVsini, has_vsini, config, phase
        self.peel, self.threaded = peel, threaded

    def lap(self, name, t0):
        import time
        now = time.time()
        self.phase[name] = self.phase.get(name, 0.0) + now - t0
        return now

    def enough(self, nwait, nstage):
        return self.peel and nwait >= max(PEEL_MIN, PEEL_FRAC * nstage)


def fit_steps(sel, fobj, specParams, fixParam, fitVsini, has_vsini, config, phase, peel=True,
              threaded=False):
    """vel_fit.process (reference vel_fit.py:505-737) for the objects `sel` of
    the engine as a coroutine: it yields requests
        ('scan', objects, grids, params, vsini)   find_best on ragged RV grids
        ('f', objects, X)                         chisq_func at fitted vectors X
        ('f0', objects, vel, vsini, params)       chisq_func0 (priors + -2 log L)
        ('model', objects, vel, params, vsini)    get_chisq(full_output=True)
        ('spawn', coroutine)                      run this coroutine beside me
    is sent their answers, and returns a list of dicts of per-object arrays (one per
    group of objects that finished together).  Every object follows the decision
    rules of the single-object path (see the module docstring).  The optimisers run
    in lock-step over the objects of a stage; with `peel` the objects that have
    finished a stage move on in groups (a spawned coroutine each) while the slower
    ones keep iterating, so the latency-bound tail of one stage runs beside the
    large calls of the next instead of in front of them."""
    import time
    t0 = time.time()
    ctx = _FitCtx(sel, fobj, specParams, fixParam, fitVsini, has_vsini, config, phase, peel,
                  threaded)
    n = len(sel)
    # 1. RV-grid scan at the starting parameters (vel_fit.py:579-602)
    vgrid = np.arange(config['min_vel'], config['max_vel'], config['vel_step0'])
    st = yield ('scan', sel, [vgrid] * n, fobj.p0[sel], fobj.vsini0[sel] if has_vsini else None)
    if not np.isfinite(st[:, :2]).all():
        raise RuntimeError('The log(likelihood) value is not finite in the initial RV scan of '
                           f'object(s) {sel[~np.isfinite(st[:, :2]).all(axis=1)].tolist()}')
    ctx.lap('scan0', t0)
    sims = simplex_starts(st[:, 1], _Sub(fobj, sel), specParams, fixParam, fitVsini,
                          config['max_vsini'])
    return (yield from _nm_stage(ctx, np.arange(n), sims, 0))


def _nm_stage_native(ctx, objs, sims, attempt, drive, stepper):
    """_nm_stage with the rounds run by the library (rvs_nm_drive): no interpreter between
    two evaluation calls.  The coroutine blocks in the library while rounds run, so it
    wants a thread of its own (run_threads)."""
    import time
    from . import _cabi
    t0 = time.time()
    eng, fobj = ctx.fobj.eng, ctx.fobj
    n = len(objs)
    handed = np.zeros(n, dtype=bool)
    try:
        while True:
            stop = int(handed.sum() + np.ceil(max(PEEL_MIN, PEEL_FRAC * n))) if ctx.peel else 0
            items0 = drive['io'].items
            rc = eng.drive_run(drive, stepper.h, SPECULATE_BELOW, stop, fobj.redo_values,
                               lambda o, X: fobj.submit(o, X)())
            fobj.nfev += drive['io'].items - items0
            if rc == _cabi.DRIVE_DONE:
                break
            act = stepper.active()
            j = np.nonzero(~act & ~handed)[0]
            res = stepper.result()
            handed[j] = True
            t0 = ctx.lap('nelder_mead', t0)
            yield ('spawn', _after_nm(ctx, objs[j], {k: v[j] for k, v in res.items()}, attempt))
        res = stepper.result()
    finally:
        eng.drive_close(drive)
        stepper.close()
    ctx.lap('nelder_mead', t0)
    j = np.nonzero(~handed)[0]
    return (yield from _after_nm(ctx, objs[j], {k: v[j] for k, v in res.items()}, attempt))


# rounds of a Nelder-Mead stage run by the library when the coroutine has its own thread
NATIVE_DRIVE = not os.environ.get('RVS_NO_NATIVE_DRIVE')


def _nm_stage(ctx, objs, sims, attempt):
    """2. Nelder-Mead for the set's objects `objs` (positions in ctx.sel), restarted once
    from its final simplex where it did not converge (vel_fit.py:628-650)."""
    import time
    t0 = time.time()
    stepper = NMStepper(sims)
    eng = ctx.fobj.eng
    if NATIVE_DRIVE and ctx.threaded and hasattr(eng, 'drive_open') and ctx.fobj.layout():
        drive = eng.drive_open(ctx.fobj.layout(), ctx.sel[objs], stepper.N, stepper.cap)
        if drive is not None:
            return (yield from _nm_stage_native(ctx, objs, sims, attempt, drive, stepper))
    handed = np.zeros(len(objs), dtype=bool)
    try:
        while True:
            req = stepper.request(SPECULATE_BELOW)
            if req is None:
                break
            stepper.feed((yield ('f', ctx.sel[objs[req[0]]], req[1])))
            if ctx.peel:
                act = stepper.active()
                j = np.nonzero(~act & ~handed)[0]
                if act.any() and ctx.enough(len(j), len(objs)):
                    res = stepper.result()
                    handed[j] = True
                    t0 = ctx.lap('nelder_mead', t0)
                    yield ('spawn', _after_nm(ctx, objs[j], {k: v[j] for k, v in res.items()},
                                              attempt))
        res = stepper.result()
    finally:
        stepper.close()
    ctx.lap('nelder_mead', t0)
    j = np.nonzero(~handed)[0]
    return (yield from _after_nm(ctx, objs[j], {k: v[j] for k, v in res.items()}, attempt))


def _after_nm(ctx, objs, res, attempt):
    out = []
    ok = res['success']
    msucc = np.ones(len(objs), dtype=bool)
    if attempt == 0 and not ok.all():
        again = np.nonzero(~ok)[0]
        rest = _nm_stage(ctx, objs[again], res['final_simplex'][again], 1)
        if ok.any():
            yield ('spawn', rest)
        else:
            return (yield from rest)
        go = np.nonzero(ok)[0]
    else:
        msucc[~ok] = False
        go = np.arange(len(objs))
    out += (yield from _bfgs_stage(ctx, objs[go], res['x'][go], msucc[go]))
    return out


# BFGS rounds stepped by the library (csrc/bfgs_host.cpp) when the coroutine has its own
# thread: microseconds per round and no interpreter lock, matrix products in index order
# (scipy's go through BLAS, so fits agree to rounding with the numpy route, not bit for bit)
NATIVE_BFGS = not os.environ.get('RVS_NO_NATIVE_BFGS')


def _bfgs_stage_native(ctx, objs, x, msucc, drive, stepper):
    import time
    from . import _cabi
    t0 = time.time()
    eng, fobj = ctx.fobj.eng, ctx.fobj
    n = len(objs)
    handed = np.zeros(n, dtype=bool)
    try:
        while True:
            stop = int(handed.sum() + np.ceil(max(PEEL_MIN, PEEL_FRAC * n))) if ctx.peel else 0
            items0 = drive['io'].items
            rc = eng.drive_run(drive, stepper.h, 0, stop, fobj.redo_values,
                               lambda o, X: fobj.submit(o, X)(), kind='bfgs')
            fobj.nfev += drive['io'].items - items0
            if rc == _cabi.DRIVE_DONE:
                break
            act = stepper.active()
            j = np.nonzero(~act & ~handed)[0]
            res = stepper.result()
            handed[j] = True
            t0 = ctx.lap('bfgs', t0)
            yield ('spawn', _finish_stage(ctx, objs[j], res['x'][j], msucc[j]))
        res = stepper.result()
    finally:
        eng.drive_close(drive)
        stepper.close()
    ctx.lap('bfgs', t0)
    j = np.nonzero(~handed)[0]
    return (yield from _finish_stage(ctx, objs[j], res['x'][j], msucc[j]))


def _bfgs_stage(ctx, objs, x, msucc):
    """3. BFGS polish (vel_fit.py:653-658)."""
    import time
    if not ctx.config.get('second_minimizer') or len(objs) == 0:
        return (yield from _finish_stage(ctx, objs, x, msucc))
    t0 = time.time()
    names = ['vel'] + (['vsini'] if ctx.fitVsini else []) + \
        [p for p in ctx.specParams if p not in ctx.fixParam]
    eng = ctx.fobj.eng
    if NATIVE_BFGS and NATIVE_DRIVE and ctx.threaded and hasattr(eng, 'drive_open') and \
            ctx.fobj.layout() and x.shape[1] <= 22:
        stepper = batch_bfgs.BFGSStepper(x, vel_fit.get_hess_inv(names))
        drive = eng.drive_open(ctx.fobj.layout(), ctx.sel[objs], stepper.N, stepper.cap)
        if drive is not None:
            return (yield from _bfgs_stage_native(ctx, objs, x, msucc, drive, stepper))
        stepper.close()
    progress = {}
    gen = batch_bfgs.bfgs_steps(x, vel_fit.get_hess_inv(names), progress=progress)
    handed = np.zeros(len(objs), dtype=bool)
    try:
        idx, X = next(gen)
        while True:
            val = yield ('f', ctx.sel[objs[idx]], X)
            idx, X = gen.send(val)
            if ctx.peel:
                fin = ~handed
                fin[progress['live']] = False
                j = np.nonzero(fin)[0]
                if ctx.enough(len(j), len(objs)):
                    handed[j] = True
                    t0 = ctx.lap('bfgs', t0)
                    yield ('spawn', _finish_stage(ctx, objs[j], progress['x'][j].copy(), msucc[j]))
    except StopIteration as stop:
        x = stop.value['x']
    ctx.lap('bfgs', t0)
    j = np.nonzero(~handed)[0]
    return (yield from _finish_stage(ctx, objs[j], x[j], msucc[j]))


def _finish_stage(ctx, objs, x, minimize_success):
    """4.-6. of vel_fit.process for objects whose optimisers have stopped at x."""
    import time
    t0 = time.time()
    n = len(objs)
    if n == 0:
        return []
    sel, fobj, config = ctx.sel[objs], ctx.fobj, ctx.config
    specParams = ctx.specParams
    loc = np.arange(n)
    min_vel, max_vel = config['min_vel'], config['max_vel']
    vel_step0, min_vel_step = config['vel_step0'], config['min_vel_step']
    vel, vsini, params, _ = fobj.unpack(sel, x)
    # 4. velocity posterior on shrinking grids (vel_fit.py:315-439), all objects per round
    best_vel = np.clip(vel, min_vel, max_vel)
    lo, hi = np.full(n, float(min_vel)), np.full(n, float(max_vel))
    step = np.full(n, float(vel_step0))
    vstat = np.zeros((n, 5))
    active = loc
    for _ in range(10):
        grids = [np.arange(np.ceil((lo[i] - best_vel[i]) / step[i]) * step[i],
                           hi[i] - best_vel[i], step[i]) + best_vel[i] for i in active]
        s = yield ('scan', sel[active], grids, params[active],
                   None if vsini is None else vsini[active])
        if not np.isfinite(s[:, :3]).all():
            raise RuntimeError('RV refinement scan without a finite minimum for object(s) '
                               f'{sel[active][~np.isfinite(s[:, :3]).all(axis=1)].tolist()}')
        vstat[active] = s
        best_vel[active] = s[:, 1]
        err = s[:, 2]
        stp = step[active]
        done = (stp < err / 5) | (stp < min_vel_step)
        coarse = stp > err
        new_step = np.where(coarse, stp / 5, err / 5 * 0.8)
        width = np.where(coarse, stp * 10, err * 10)
        cont = ~done
        a2 = active[cont]
        lo[a2] = np.maximum(best_vel[a2] - width[cont], lo[a2])
        hi[a2] = np.minimum(best_vel[a2] + width[cont], hi[a2])
        step[a2] = new_step[cont]
        active = a2
        if len(active) == 0:
            break
    t0 = ctx.lap('refine', t0)
    # 5. model at the best point (vel_fit.py:688-696)
    tot, info = yield ('model', sel, best_vel, params, vsini)
    if not np.isfinite(tot).all():
        raise RuntimeError('The log(likelihood) value is not finite at the best-fit point of '
                           f'object(s) {sel[~np.isfinite(tot)].tolist()}')
    t0 = ctx.lap('model', t0)
    # 6. Hessian over the atmospheric parameters at the optimiser's velocity and vsini
    #    (vel_fit.py:698-722: hess_func keeps best_param's own velocity)
    hsteps = [vel_fit.HESS_STEP[_] for _ in specParams]
    P = hessian_points(params, hsteps)
    npts = P.shape[1]
    ii = np.repeat(loc, npts)
    vals = yield ('f0', sel[ii], vel[ii], None if vsini is None else vsini[ii],
                  P.reshape(n * npts, -1))
    hess = hessian_from_values(0.5 * vals.reshape(n, npts), hsteps)
    ctx.lap('hessian', t0)
    return [dict(sel=sel, x=x, vel=vel, vsini=vsini, params=params, best_vel=best_vel,
                 vstat=vstat, minimize_success=minimize_success, chisq=tot, info=info,
                 hessian=hess)]


class _Sub:
    """The rows `sel` of a BatchObjective's per-object start values."""

    def __init__(self, fobj, sel):
        self.p0 = fobj.p0[sel]
        self.vsini0 = fobj.vsini0[sel]


# objects per lock-step set and sets in flight: enough sets that the tail of one (few
# live problems, latency-bound calls) runs under the bulk of the others
FIT_GROUP = 256
FIT_MAX_GROUPS = 2
FIT_MAX_SET = 4096       # objects per set at most (auto grouping)
THREADS = True
PEEL = True
FIT_SPLIT = {2: [3, 2], 3: [5, 4, 3], 4: [4, 3, 2, 1]}      # number of sets -> relative sizes (else equal)


def process_batch(objects, paramDict0s, fixParam=None, options=None, config=None, priors=None,
                  engine=None, timer=None, groups=None, threads=None, peel=None):
    """vel_fit.process for a list of objects (each a list of SpecData); same
    arguments otherwise, paramDict0s one dictionary per object.  Returns the list
    of result dictionaries of vel_fit.process.  `engine`: a LikelihoodEngine
    already holding the objects on the device (then `objects` may be None).
    The objects are fitted as `groups` independent lock-step sets whose phases
    interleave on the GPU: one host thread per set (run_threads; `threads=False`:
    all sets advanced by the calling thread, run_pipeline).  `peel` (default: on when the
    sets are large enough to be worth splitting): objects that have finished an optimiser
    stage move on in groups while the slower ones keep iterating (see fit_steps)."""
    if config is None:
        raise RuntimeError('Config must be provided')
    options = options or {}
    fixParam = fixParam or []
    if engine is not None:
        objects = engine.objects
    objects = [[o] if isinstance(o, spec_fit.SpecData) else list(o) for o in objects]
    B = len(objects)
    eng = engine if engine is not None else spec_fit.LikelihoodEngine(objects, config, options)
    eng.timer = timer
    setup0 = objects[0][0].name
    specParams = list(spec_inter.getSpecParams(setup0, config))
    has_vsini = 'vsini' in paramDict0s[0]
    fitVsini = has_vsini and 'vsini' not in fixParam
    fobj = BatchObjective(eng, specParams, paramDict0s, fixParam, fitVsini, config, priors)
    fobj.layout()       # built once, before the sets' threads ask for it
    if groups is None:
        # a few sets whose phases interleave; more of them when the batch is so large that
        # the first request of a set (its whole simplices) would make the per-slot device
        # buffers huge
        groups = int(max(np.clip(B // FIT_GROUP, 1, FIT_MAX_GROUPS), -(-B // FIT_MAX_SET)))
    # a set holds one evaluation slot through its Nelder-Mead stage and needs up to two
    # more for a request that goes out in pieces
    groups = max(1, min(groups, B, eng.NSLOT // 2))
    phase = {}
    # sets of unequal size reach their latency-bound stretches (optimiser tails, refinement
    # scans, model output) at different times, under the large calls of the others
    frac = np.array(FIT_SPLIT.get(groups) or [1.0 / groups] * groups, dtype=np.float64)
    cuts = np.round(np.cumsum(frac / frac.sum()) * B).astype(int)[:-1]
    parts = [p for p in np.split(np.arange(B), cuts) if len(p)]
    if peel is None:
        peel = PEEL and hasattr(eng, 'submit_fit')
    if threads is None:
        threads = THREADS and hasattr(eng, 'submit_fit')
    gens = [fit_steps(p, fobj, specParams, fixParam, fitVsini, has_vsini, config, phase, peel,
                      bool(threads)) for p in parts]

    def start(req):
        kind = req[0]
        if kind == 'f':
            return fobj.submit(req[1], req[2])
        if kind == 'f0':
            return fobj.submit0(*req[1:])
        if kind == 'scan':
            return _Ready(_scan_round(eng, *req[1:]))
        if kind == 'model':
            _, idx, vel, params, vsini = req
            with general_lock:
                return _Ready(eng.evaluate(idx, vel[:, None], params, vsini, want_model=True))
        raise ValueError(kind)
    def assemble(parts_done):
        """Result dictionaries of the groups a coroutine finished: [(object index, dict)]."""
        rows_out = []
        for r in parts_done:
            rows_out += assemble_one(r)
        return rows_out

    def assemble_one(r):
        rows_out = []
        info = r['info']
        tot = r['chisq'][:, 0]
        rows = {name: {int(o): j for j, o in enumerate(arm['sel'])}
                for name, arm in info['arms'].items()}
        for k, i in enumerate(r['sel']):
            diag_err, covar, bad_hessian = vel_fit._uncertainties_from_hessian(r['hessian'][k])
            ret = dict(param=dict(zip(specParams, r['params'][k])))
            if fitVsini:
                ret['vsini'] = r['vsini'][k]
            ret.update(vel=r['best_vel'][k], vel_err=r['vstat'][k, 2],
                       vel_skewness=r['vstat'][k, 3], vel_kurtosis=r['vstat'][k, 4],
                       param_err=dict(zip(specParams, diag_err)), param_covar=covar,
                       minimize_success=bool(r['minimize_success'][k]),
                       bad_hessian=bad_hessian, chisq=float(tot[k]), logl=-0.5 * float(tot[k]),
                       yfit=[], raw_models=[], chisq_array=[], npix_array=[])
            for sd in objects[i]:
                arm = info['arms'][sd.name]
                j = rows[sd.name][k]
                if arm['tbad'][j]:
                    ret['chisq_array'].append(np.nan)
                    ret['yfit'].append(np.zeros(len(sd.lam)) + np.nan)
                    continue
                ex = arm['extras']
                sl = slice(ex['moff'][j], ex['moff'][j + 1])
                model, raw = ex['model'][sl], ex['raw'][sl]
                good = ~sd.badmask
                ret['yfit'].append(model)
                ret['raw_models'].append(raw)
                ret['chisq_array'].append(float(np.sum((((model - sd.spec) / sd.espec)[good])**2)))
                ret['npix_array'].append(int(good.sum()))
            rows_out.append((int(i), ret))
        return rows_out
    import contextlib
    general_lock = getattr(eng, '_general_lock', contextlib.nullcontext())
    try:
        if threads:
            device = _dev.torch_mod().cuda.current_device() if hasattr(eng, 'submit_fit') \
                else None
            results = run_threads(gens, start, device, post=assemble)
        else:
            results = [assemble(r) for r in run_pipeline(gens, start)]
    finally:
        eng.timer = None
        eng.drain()
    out = [None] * B
    for rows in results:
        for i, ret in rows:
            out[i] = ret
    process_batch.last_phase_seconds = phase
    process_batch.last_spawned = len(results) - len(parts)
    return out
