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
  * BFGS *is* scipy's `minimize(method='BFGS')`, one instance per object on its
    own thread; every function value and every finite-difference gradient
    (scipy's `workers` hook of approx_derivative) blocks on a coordinator that
    gathers the requests of all live objects into one launch;
  * the Hessian is the same central-difference routine as vel_fit, its
    evaluation points recorded, evaluated in one launch and replayed.
"""
import threading

import numpy as np
import scipy.optimize

from . import _dev, spec_fit, spec_inter, vel_fit


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
        self.nfev += len(idx)
        chi = self.eng.evaluate(idx, vel, params, vsini)
        return self.prior_term(params) + chi

    def __call__(self, idx, X):
        """chisq_func for K pairs: idx (K,) object indices, X (K, N)."""
        return self.submit(idx, X)()

    def submit(self, idx, X):
        """Start the evaluation of chisq_func for K pairs and return a callable that
        waits for the values (same arithmetic as __call__: priors + -2 log L +
        penalty, 1e30 behind the hard walls, vel_fit.py:210-257)."""
        idx = np.asarray(idx, dtype=np.int64)
        vel, vsini, params, pen = self.unpack(idx, X)
        wall = (vel > self.max_vel) | (vel < self.min_vel) | ~np.isfinite(params).all(axis=1)
        ok = ~wall
        pend = None
        if ok.any():
            self.nfev += int(ok.sum())
            if hasattr(self.eng, 'submit'):
                pend = self.eng.submit(idx[ok], vel[ok], params[ok],
                                       None if vsini is None else vsini[ok])
            else:       # engines with a blocking evaluate only
                class _Done:
                    def __init__(self, v):
                        self.v = v

                    def result(self):
                        return self.v
                pend = _Done(self.eng.evaluate(idx[ok], vel[ok], params[ok],
                                               None if vsini is None else vsini[ok]))

        def wait():
            out = np.full(len(idx), 1e30)
            if pend is not None:
                out[ok] = self.prior_term(params[ok]) + pend.result() + pen[ok]
            return out
        return wait


# below this many active problems an iteration's candidate points go out in one call
SPECULATE_BELOW = 640
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


def nelder_mead_lockstep(fbatch, sims, xatol=1e-2, fatol=1e-3, maxiter=10000,
                         speculate_below=0):
    """nelder_mead_steps driven with a blocking objective fbatch(idx, X) -> f."""
    gen = nelder_mead_steps(sims, xatol, fatol, maxiter, speculate_below)
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


# ------------------------------------------------- scipy BFGS, many at a time
class _Coordinator:
    """Gathers the blocking evaluation requests of many worker threads into one
    batched call.  A round is evaluated when every live worker has a request
    pending."""

    def __init__(self, fbatch, ids):
        self.fbatch = fbatch
        self.lock = threading.Lock()
        self.ready = threading.Event()
        self.events = {i: threading.Event() for i in ids}
        self.pending, self.results = {}, {}
        self.live = len(ids)
        self.rounds = 0

    def request(self, wid, X):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        with self.lock:
            self.pending[wid] = X
            if len(self.pending) >= self.live:
                self.ready.set()
        ev = self.events[wid]
        ev.wait()
        ev.clear()
        return self.results.pop(wid)

    def finished(self, wid):
        with self.lock:
            self.live -= 1
            if self.live == 0 or len(self.pending) >= self.live:
                self.ready.set()

    def run(self):
        while True:
            self.ready.wait()
            with self.lock:
                self.ready.clear()
                if self.live == 0 and not self.pending:
                    return
                if len(self.pending) < self.live:
                    continue
                batch, self.pending = self.pending, {}
            wids = list(batch)
            idx = np.concatenate([np.full(len(batch[w]), w) for w in wids])
            vals = self.fbatch(idx, np.concatenate([batch[w] for w in wids]))
            self.rounds += 1
            pos = 0
            for w in wids:
                n = len(batch[w])
                self.results[w] = vals[pos:pos + n]
                pos += n
                self.events[w].set()


def bfgs_batch(fbatch, x0s, hess_inv0, ids=None):
    """scipy.optimize.minimize(method='BFGS', options=dict(hess_inv0=...)) for every
    row of x0s, the instances running concurrently and sharing launches.  Returns
    the list of scipy OptimizeResult objects."""
    x0s = np.asarray(x0s, dtype=np.float64)
    ids = list(range(len(x0s))) if ids is None else list(ids)
    coord = _Coordinator(fbatch, ids)
    results, errors = {}, {}

    def worker(wid, x0):
        try:
            def fun(x):
                return float(coord.request(wid, x)[0])

            def pmap(_f, it):
                xs = [np.asarray(_) for _ in it]
                if not xs:
                    return []
                return [float(_) for _ in coord.request(wid, np.array(xs))]
            results[wid] = scipy.optimize.minimize(
                fun, x0, method='BFGS', options=dict(hess_inv0=hess_inv0, workers=pmap))
        except BaseException as exc:          # noqa: BLE001  (re-raised in the caller)
            errors[wid] = exc
        finally:
            coord.finished(wid)

    old = threading.stack_size(512 * 1024)
    try:
        threads = [threading.Thread(target=worker, args=(w, x0s[i]), daemon=True)
                   for i, w in enumerate(ids)]
        for t in threads:
            t.start()
    finally:
        threading.stack_size(old)
    coord.run()
    for t in threads:
        t.join()
    if errors:
        raise next(iter(errors.values()))
    return [results[w] for w in ids]


# ---- the same, with the per-object scipy drivers spread over worker processes
def _bfgs_process_main(conn):
    """Worker process: runs scipy BFGS for a subset of the objects (threads +
    coordinator as above); every evaluation round is one message to the parent,
    which owns the GPU.  Never touches CUDA."""
    while True:
        try:
            job = conn.recv()
        except EOFError:
            return
        if job is None:
            return
        ids, x0s, hess_inv0 = job

        def remote(idx, X):
            conn.send(('req', np.asarray(idx), np.asarray(X)))
            return conn.recv()
        try:
            res = bfgs_batch(remote, x0s, hess_inv0, ids=ids)
            conn.send(('done', {w: dict(r) for w, r in zip(ids, res)}))
        except BaseException as exc:      # noqa: BLE001
            conn.send(('error', repr(exc)))


class _BfgsProcessPool:
    def __init__(self, nproc):
        import multiprocessing as mp
        ctx = mp.get_context('spawn')
        self.conns, self.procs = [], []
        for _ in range(nproc):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_bfgs_process_main, args=(child,), daemon=True)
            p.start()
            child.close()
            self.conns.append(parent)
            self.procs.append(p)

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except (OSError, ValueError):
                pass
        for p in self.procs:
            p.join(timeout=2)

    def run(self, fbatch, x0s, hess_inv0):
        B = len(x0s)
        chunks = [c for c in np.array_split(np.arange(B), len(self.conns)) if len(c)]
        live = {}
        for conn, ch in zip(self.conns, chunks):
            conn.send((list(map(int, ch)), x0s[ch], hess_inv0))
            live[conn] = True
        results = {}
        while live:
            reqs = []
            for conn in list(live):
                msg = conn.recv()
                if msg[0] == 'req':
                    reqs.append((conn, msg[1], msg[2]))
                elif msg[0] == 'done':
                    results.update(msg[1])
                    del live[conn]
                else:
                    raise RuntimeError('BFGS worker failed: ' + msg[1])
            if reqs:
                vals = fbatch(np.concatenate([r[1] for r in reqs]),
                              np.concatenate([r[2] for r in reqs]))
                pos = 0
                for conn, idx, _ in reqs:
                    conn.send(vals[pos:pos + len(idx)])
                    pos += len(idx)
        return [scipy.optimize.OptimizeResult(results[i]) for i in range(B)]


_bfgs_pool = None


def bfgs_many(fbatch, x0s, hess_inv0, nproc=None):
    """bfgs_batch, with the scipy drivers (pure Python, GIL-bound) spread over a
    persistent pool of worker processes when there are enough objects."""
    import atexit
    import os
    global _bfgs_pool
    x0s = np.asarray(x0s, dtype=np.float64)
    if nproc is None:
        nproc = min(os.cpu_count() or 1, 16, len(x0s) // 32)
    if nproc < 2:
        return bfgs_batch(fbatch, x0s, hess_inv0)
    if _bfgs_pool is None or len(_bfgs_pool.conns) != nproc:
        if _bfgs_pool is not None:
            _bfgs_pool.close()
        _bfgs_pool = _BfgsProcessPool(nproc)
        atexit.register(_bfgs_pool.close)
    return _bfgs_pool.run(fbatch, x0s, hess_inv0)


# ------------------------------------------------------------ the batched fit
class _Replay:
    """Records the points a routine asks for, then replays their values."""

    def __init__(self):
        self.pts, self.vals, self.i = [], None, 0

    def __call__(self, x):
        if self.vals is None:
            self.pts.append(np.array(x, dtype=np.float64))
            return 0.0
        v = self.vals[self.i]
        self.i += 1
        return v


def _scan_round(eng, idx, grids, params, vsini):
    """find_best (one template per object) for ragged velocity grids: chi-squares
    in one launch per arm, statistics on the device per distinct grid length.
    Returns (n, 5): best_chi, best_vel, vel_err, skewness, kurtosis."""
    nv = np.array([len(g) for g in grids])
    V = np.stack([np.concatenate([g, np.full(nv.max() - len(g), g[-1])]) for g in grids])
    chi = eng.evaluate(idx, V, params, vsini)
    out = np.zeros((len(idx), 5))
    for n in np.unique(nv):
        sel = np.nonzero(nv == n)[0]
        st, _ = spec_fit.scan_stats(V[sel, :n], chi[sel, None, :n])
        out[sel] = st[:, :5]
    return out


def process_batch(objects, paramDict0s, fixParam=None, options=None, config=None, priors=None,
                  engine=None, timer=None):
    """vel_fit.process for a list of objects (each a list of SpecData); same
    arguments otherwise, paramDict0s one dictionary per object.  Returns the list
    of result dictionaries of vel_fit.process.  `engine`: a LikelihoodEngine
    already holding the objects on the device (then `objects` may be None)."""
    if config is None:
        raise RuntimeError('Config must be provided')
    options = options or {}
    fixParam = fixParam or []
    if engine is not None:
        objects = engine.objects
    objects = [[o] if isinstance(o, spec_fit.SpecData) else list(o) for o in objects]
    B = len(objects)
    min_vel, max_vel = config['min_vel'], config['max_vel']
    vel_step0, min_vel_step = config['vel_step0'], config['min_vel_step']
    second_minimizer = config.get('second_minimizer') or False
    eng = engine if engine is not None else spec_fit.LikelihoodEngine(objects, config, options)
    eng.timer = timer
    import time
    phase = {}
    t_last = [time.time()]

    def lap(name):
        now = time.time()
        phase[name] = phase.get(name, 0.0) + now - t_last[0]
        t_last[0] = now
    setup0 = objects[0][0].name
    specParams = list(spec_inter.getSpecParams(setup0, config))
    has_vsini = 'vsini' in paramDict0s[0]
    fitVsini = has_vsini and 'vsini' not in fixParam
    vsiniMapper = vel_fit.VSiniMapper(config['max_vsini']) if fitVsini else None
    fobj = BatchObjective(eng, specParams, paramDict0s, fixParam, fitVsini, config, priors)
    allb = np.arange(B)
    # 1. RV-grid scan at the starting parameters (vel_fit.py:579-590)
    vgrid = np.arange(min_vel, max_vel, vel_step0)
    st = _scan_round(eng, allb, [vgrid] * B, fobj.p0,
                     fobj.vsini0 if has_vsini else None)
    lap('scan0')
    # 2. Nelder-Mead, restarted once from its final simplex if it did not converge
    sims = np.stack([vel_fit._get_simplex_start(
        st[i, 1], fixParam=fixParam, specParamNames=specParams, paramDict0=paramDict0s[i],
        vsiniMapper=vsiniMapper, fitVsini=fitVsini)[1] for i in range(B)])
    minimize_success = np.ones(B, dtype=bool)
    x = np.zeros((B, sims.shape[2]))
    todo = allb
    for attempt in range(2):
        # two independent lock-step sets: the host advances one simplex set while the
        # other's trial points are on the GPU
        res = nelder_mead_interleaved(lambda i, X: fobj.submit(todo[i], X), sims[todo],
                                      groups=2 if len(todo) >= 2 * NM_MIN_GROUP else 1,
                                      speculate_below=SPECULATE_BELOW)
        x[todo] = res['x']
        sims[todo] = res['final_simplex']
        failed = todo[~res['success']]
        if attempt == 1:
            minimize_success[failed] = False
        todo = failed
        if len(todo) == 0:
            break
    lap('nelder_mead')
    # 3. BFGS polish (vel_fit.py:653-658)
    if second_minimizer:
        names = ['vel'] + (['vsini'] if fitVsini else []) + \
            [p for p in specParams if p not in fixParam]
        bres = bfgs_many(fobj, x, vel_fit.get_hess_inv(names))
        x = np.array([r['x'] for r in bres])
    lap('bfgs')
    vel, vsini, params, _ = fobj.unpack(allb, x)
    # 4. velocity posterior on shrinking grids (vel_fit.py:315-439), all objects per round
    best_vel = np.clip(vel, min_vel, max_vel)
    lo, hi = np.full(B, float(min_vel)), np.full(B, float(max_vel))
    step = np.full(B, float(vel_step0))
    vstat = np.zeros((B, 5))
    active = allb
    for _ in range(10):
        grids = [np.arange(np.ceil((lo[i] - best_vel[i]) / step[i]) * step[i],
                           hi[i] - best_vel[i], step[i]) + best_vel[i] for i in active]
        s = _scan_round(eng, active, grids, params[active],
                        None if vsini is None else vsini[active])
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
    lap('refine')
    # 5. model at the best point (vel_fit.py:688-696)
    tot, info = eng.evaluate(allb, best_vel[:, None], params, vsini, want_model=True)
    chi_best = tot[:, 0]
    lap('model')
    # 6. Hessian over the atmospheric parameters at the optimiser's velocity and vsini
    #    (vel_fit.py:698-722: hess_func keeps best_param's own velocity)
    hsteps = [vel_fit.HESS_STEP[_] for _ in specParams]
    recs = []
    for i in range(B):
        r = _Replay()
        vel_fit.central_hessian(r, params[i], hsteps)
        recs.append(r)
    npts = [len(r.pts) for r in recs]
    P = np.concatenate([np.array(r.pts) for r in recs])
    ii = np.repeat(allb, npts)
    vals = 0.5 * fobj.chisq0(ii, vel[ii], None if vsini is None else vsini[ii], P)
    out, pos = [], 0
    for i in range(B):
        recs[i].vals = vals[pos:pos + npts[i]]
        pos += npts[i]
        hessian = vel_fit.central_hessian(recs[i], params[i], hsteps)
        diag_err, covar, bad_hessian = vel_fit._uncertainties_from_hessian(hessian)
        ret = dict(param=dict(zip(specParams, params[i])))
        if fitVsini:
            ret['vsini'] = vsini[i]
        ret.update(vel=best_vel[i], vel_err=vstat[i, 2], vel_skewness=vstat[i, 3],
                   vel_kurtosis=vstat[i, 4], param_err=dict(zip(specParams, diag_err)),
                   param_covar=covar, minimize_success=bool(minimize_success[i]),
                   bad_hessian=bad_hessian, chisq=float(chi_best[i]),
                   logl=-0.5 * float(chi_best[i]), yfit=[], raw_models=[], chisq_array=[],
                   npix_array=[])
        for sd in objects[i]:
            arm = info['arms'][sd.name]
            j = int(np.nonzero(arm['sel'] == i)[0][0])
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
        out.append(ret)
    lap('hessian')
    eng.timer = None
    process_batch.last_phase_seconds = phase
    return out
