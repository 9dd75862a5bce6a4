"""scipy's BFGS for many problems in lock-step.

`vel_fit.process` polishes the Nelder-Mead optimum with
scipy.optimize.minimize(method='BFGS', options=dict(hess_inv0=...)) and a
forward-difference gradient (reference vel_fit.py:653-658).  One such
minimisation asks for its function values one after the other; a thousand of
them ask for a thousand values at a time.  `bfgs_steps` is that algorithm --
`_minimize_bfgs` with the MINPACK line search `DCSRCH` (scipy/optimize/
_optimize.py, _linesearch.py, _dcsrch.py, _numdiff.py; scipy 1.16+), the same
constants, the same tests in the same order, the same floating-point
expressions -- written over arrays of problems: every round needs one function
value and one forward-difference gradient per live problem (N + 1 points each),
which the caller evaluates in ONE launch.  The decision logic between rounds is
numpy over the live problems.  The fallback line search (`line_search_wolfe2`,
entered once per problem at most, typically when the finite-difference gradient
can no longer deliver a descent step) runs as a small per-problem coroutine.

Problem b visits exactly the points scipy would visit and returns exactly
scipy's x, fun, nit, status (tests/test_batch_drivers.py compares bit for bit).
"""
import numpy as np

_EPSILON = np.sqrt(np.finfo(float).eps)      # scipy.optimize._optimize._epsilon

# DCSRCH constants (_dcsrch.py)
_P5, _P66, _XTRAPL, _XTRAPU = 0.5, 0.66, 1.1, 4.0
_FG, _CONV, _WARN, _ERROR = 0, 1, 2, 3


def _pymax(a, b):
    """Python's max(a, b) element-wise (b only if b > a: NaNs do not propagate
    the way np.maximum's do)."""
    return np.where(b > a, b, a)


def _pymin(a, b):
    return np.where(b < a, b, a)


def fd_points(x, eps=_EPSILON):
    """Rows scipy's 2-point `approx_derivative(abs_step=eps)` evaluates for the
    gradient at every row of x (m, N), preceded by x itself: (m, N + 1, N), and
    the denominators (m, N) (_numdiff.py:585-597, 694-711)."""
    m, N = x.shape
    h = np.full((m, N), eps)
    dx = (x + h) - x
    sign = (x >= 0).astype(np.float64) * 2 - 1
    h = np.where(dx == 0, _EPSILON * sign * np.maximum(1.0, np.abs(x)), h)
    P = np.repeat(x[:, None, :], N + 1, axis=1)
    i = np.arange(N)
    P[:, 1 + i, i] = x[:, i] + h[:, i]
    return P, (x + h) - x


def _dcstep(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stpmin, stpmax):
    """MINPACK-2 dcstep (_dcsrch.py), all four cases computed and selected."""
    sgnd = np.sign(dp) * np.sign(dx)
    with np.errstate(all='ignore'):
        c1 = fp > fx
        c2 = ~c1 & (sgnd < 0.0)
        c3 = ~c1 & ~c2 & (np.abs(dp) < np.abs(dx))
        # case 1: higher function value, the minimum is bracketed
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = _pymax(_pymax(np.abs(theta), np.abs(dx)), np.abs(dp))
        gamma = s * np.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        g1 = np.where(stp < stx, -gamma, gamma)
        p = (g1 - dx) + theta
        q = ((g1 - dx) + g1) + dp
        r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        f1 = np.where(np.abs(stpc - stx) <= np.abs(stpq - stx), stpc, stpc + (stpq - stpc) / 2.0)
        # case 2: lower value, derivatives of opposite sign
        g2 = np.where(stp > stx, -gamma, gamma)
        p = (g2 - dp) + theta
        q = ((g2 - dp) + g2) + dx
        r = p / q
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        f2 = np.where(np.abs(stpc - stp) > np.abs(stpq - stp), stpc, stpq)
        # case 3: lower value, same sign, the derivative decreases
        rad = (theta / s) ** 2 - (dx / s) * (dp / s)
        gamma3 = s * np.sqrt(np.where(rad > 0, rad, 0.0))
        g3 = np.where(stp > stx, -gamma3, gamma3)
        p = (g3 - dp) + theta
        q = (g3 + (dx - dp)) + g3
        r = p / q
        stpc = np.where((r < 0) & (g3 != 0), stp + r * (stx - stp),
                        np.where(stp > stx, stpmax, stpmin))
        lim = stp + 0.66 * (sty - stp)
        f3b = np.where(np.abs(stpc - stp) < np.abs(stpq - stp), stpc, stpq)
        f3b = np.where(stp > stx, _pymin(lim, f3b), _pymax(lim, f3b))
        f3n = np.clip(np.where(np.abs(stpc - stp) > np.abs(stpq - stp), stpc, stpq),
                      stpmin, stpmax)
        f3 = np.where(brackt, f3b, f3n)
        # case 4: lower value, same sign, the derivative does not decrease
        theta4 = 3.0 * (fp - fy) / (sty - stp) + dy + dp
        s4 = _pymax(_pymax(np.abs(theta4), np.abs(dy)), np.abs(dp))
        gamma4 = s4 * np.sqrt((theta4 / s4) ** 2 - (dy / s4) * (dp / s4))
        g4 = np.where(stp > sty, -gamma4, gamma4)
        p = (g4 - dp) + theta4
        q = ((g4 - dp) + g4) + dy
        r = p / q
        f4 = np.where(brackt, stp + r * (sty - stp), np.where(stp > stx, stpmax, stpmin))
    stpf = np.where(c1, f1, np.where(c2, f2, np.where(c3, f3, f4)))
    brackt = brackt | c1 | c2
    swap = ~c1 & (sgnd < 0)
    nsty = np.where(c1, stp, np.where(swap, stx, sty))
    nfy = np.where(c1, fp, np.where(swap, fx, fy))
    ndy = np.where(c1, dp, np.where(swap, dx, dy))
    nstx = np.where(c1, stx, stp)
    nfx = np.where(c1, fx, fp)
    ndx = np.where(c1, dx, dp)
    return nstx, nfx, ndx, nsty, nfy, ndy, stpf, brackt


class _Dcsrch:
    """State of DCSRCH (_dcsrch.py) for B line searches; methods act on the
    subset `s` (index array)."""

    NAMES = ('finit', 'ginit', 'gtest', 'width', 'width1', 'stx', 'fx', 'gx', 'sty', 'fy',
             'gy', 'stmin', 'stmax')

    def __init__(self, B, ftol, gtol, xtol, stpmin, stpmax):
        for k in self.NAMES:
            setattr(self, k, np.zeros(B))
        self.brackt = np.zeros(B, dtype=bool)
        self.stage = np.ones(B, dtype=np.int64)
        self.ftol, self.gtol, self.xtol = ftol, gtol, xtol
        self.stpmin, self.stpmax = stpmin, stpmax

    def start(self, s, stp, f, g):
        """task START: returns task (FG or ERROR) per problem of s."""
        err = (stp < self.stpmin) | (stp > self.stpmax) | (g >= 0)
        self.brackt[s] = False
        self.stage[s] = 1
        self.finit[s], self.ginit[s] = f, g
        self.gtest[s] = self.ftol * g
        self.width[s] = self.stpmax - self.stpmin
        self.width1[s] = (self.stpmax - self.stpmin) / _P5
        self.stx[s], self.fx[s], self.gx[s] = 0.0, f, g
        self.sty[s], self.fy[s], self.gy[s] = 0.0, f, g
        self.stmin[s] = 0.0
        self.stmax[s] = stp + _XTRAPU * stp
        return np.where(err, _ERROR, _FG)

    def iterate(self, s, stp, f, g):
        """One call of DCSRCH._iterate with task FG: (new stp, task)."""
        finit, ginit, gtest = self.finit[s], self.ginit[s], self.gtest[s]
        brackt, stage = self.brackt[s], self.stage[s]
        stmin, stmax = self.stmin[s], self.stmax[s]
        stx, fx, gx = self.stx[s], self.fx[s], self.gx[s]
        sty, fy, gy = self.sty[s], self.fy[s], self.gy[s]
        with np.errstate(all='ignore'):
            ftest = finit + stp * gtest
            stage = np.where((stage == 1) & (f <= ftest) & (g >= 0), 2, stage)
            warn = brackt & ((stp <= stmin) | (stp >= stmax))
            warn |= brackt & (stmax - stmin <= self.xtol * stmax)
            warn |= (stp == self.stpmax) & (f <= ftest) & (g <= gtest)
            warn |= (stp == self.stpmin) & ((f > ftest) | (g >= gtest))
            conv = (f <= ftest) & (np.abs(g) <= self.gtol * -ginit)
            task = np.where(conv, _CONV, np.where(warn, _WARN, _FG))
            go = task == _FG
            # the modified function of stage 1
            mod = (stage == 1) & (f <= fx) & (f > ftest)
            a_f = np.where(mod, f - stp * gtest, f)
            a_fx = np.where(mod, fx - stx * gtest, fx)
            a_fy = np.where(mod, fy - sty * gtest, fy)
            a_g = np.where(mod, g - gtest, g)
            a_gx = np.where(mod, gx - gtest, gx)
            a_gy = np.where(mod, gy - gtest, gy)
            nstx, nfx, ngx, nsty, nfy, ngy, nstp, nbr = _dcstep(
                stx, a_fx, a_gx, sty, a_fy, a_gy, stp, a_f, a_g, brackt, stmin, stmax)
            nfx = np.where(mod, nfx + nstx * gtest, nfx)
            nfy = np.where(mod, nfy + nsty * gtest, nfy)
            ngx = np.where(mod, ngx + gtest, ngx)
            ngy = np.where(mod, ngy + gtest, ngy)
            width, width1 = self.width[s], self.width1[s]
            bis = nbr & (np.abs(nsty - nstx) >= _P66 * width1)
            nstp = np.where(bis, nstx + _P5 * (nsty - nstx), nstp)
            width1 = np.where(nbr, width, width1)
            width = np.where(nbr, np.abs(nsty - nstx), width)
            nstmin = np.where(nbr, _pymin(nstx, nsty), nstp + _XTRAPL * (nstp - nstx))
            nstmax = np.where(nbr, _pymax(nstx, nsty), nstp + _XTRAPU * (nstp - nstx))
            nstp = np.clip(nstp, self.stpmin, self.stpmax)
            back = (nbr & ((nstp <= nstmin) | (nstp >= nstmax))) | \
                (nbr & (nstmax - nstmin <= self.xtol * nstmax))
            nstp = np.where(back, nstx, nstp)
        # problems that stopped (CONV / WARN) returned before touching the bracket
        self.stage[s] = stage
        keep = lambda new, old: np.where(go, new, old)      # noqa: E731
        self.brackt[s] = keep(nbr, brackt)
        self.stx[s], self.fx[s], self.gx[s] = keep(nstx, stx), keep(nfx, fx), keep(ngx, gx)
        self.sty[s], self.fy[s], self.gy[s] = keep(nsty, sty), keep(nfy, fy), keep(ngy, gy)
        self.width[s], self.width1[s] = keep(width, self.width[s]), keep(width1, self.width1[s])
        self.stmin[s], self.stmax[s] = keep(nstmin, stmin), keep(nstmax, stmax)
        return np.where(go, nstp, stp), task


# ------------------------------------------------------- fallback line search
def _cubicmin(a, fa, fpa, b, fb, c, fc):
    """Minimiser of the cubic through (a, fa), (b, fb), (c, fc) with slope fpa at
    a; None where scipy's version (_linesearch.py) gives up."""
    with np.errstate(divide='raise', over='raise', invalid='raise'):
        try:
            db, dc = b - a, c - a
            denom = (db * dc) ** 2 * (db - dc)
            d1 = np.empty((2, 2))
            d1[0, 0], d1[0, 1] = dc ** 2, -db ** 2
            d1[1, 0], d1[1, 1] = -dc ** 3, db ** 3
            A, B = np.dot(d1, np.asarray([fb - fa - fpa * db, fc - fa - fpa * dc]).flatten())
            A /= denom
            B /= denom
            xmin = a + (-B + np.sqrt(B * B - 3 * A * fpa)) / (3 * A)
        except ArithmeticError:
            return None
    return xmin if np.isfinite(xmin) else None


def _quadmin(a, fa, fpa, b, fb):
    with np.errstate(divide='raise', over='raise', invalid='raise'):
        try:
            db = b - a * 1.0
            B = (fb - fa - fpa * db) / (db * db)
            xmin = a - fpa / (2.0 * B)
        except ArithmeticError:
            return None
    return xmin if np.isfinite(xmin) else None


def _zoom(a_lo, a_hi, phi_lo, phi_hi, derphi_lo, phi0, derphi0, c1, c2):
    """scipy's _zoom as a coroutine: yields a step length, is sent (phi, derphi)
    there; returns (a_star, phi_star, derphi_star) or (None, None, None)."""
    i, phi_rec, a_rec = 0, phi0, 0
    while True:
        dalpha = a_hi - a_lo
        a, b = (a_hi, a_lo) if dalpha < 0 else (a_lo, a_hi)
        a_j = None
        if i > 0:
            cchk = 0.2 * dalpha
            a_j = _cubicmin(a_lo, phi_lo, derphi_lo, a_hi, phi_hi, a_rec, phi_rec)
        if i == 0 or a_j is None or a_j > b - cchk or a_j < a + cchk:
            qchk = 0.1 * dalpha
            a_j = _quadmin(a_lo, phi_lo, derphi_lo, a_hi, phi_hi)
            if a_j is None or a_j > b - qchk or a_j < a + qchk:
                a_j = a_lo + 0.5 * dalpha
        phi_aj, derphi_aj = yield a_j
        if phi_aj > phi0 + c1 * a_j * derphi0 or phi_aj >= phi_lo:
            phi_rec, a_rec, a_hi, phi_hi = phi_hi, a_hi, a_j, phi_aj
        else:
            if abs(derphi_aj) <= -c2 * derphi0:
                return a_j, phi_aj, derphi_aj
            if derphi_aj * (a_hi - a_lo) >= 0:
                phi_rec, a_rec, a_hi, phi_hi = phi_hi, a_hi, a_lo, phi_lo
            else:
                phi_rec, a_rec = phi_lo, a_lo
            a_lo, phi_lo, derphi_lo = a_j, phi_aj, derphi_aj
        i += 1
        if i > 10:
            return None, None, None


def _wolfe2(phi0, old_phi0, derphi0, c1, c2, amax, maxiter=10):
    """scipy's scalar_search_wolfe2 as a coroutine (every step length it asks
    for gets the function value AND the slope; scipy asks for the slope only
    after the value, at the same point, so the visited points are the same).
    Returns (alpha_star, phi_star, have_gradient): have_gradient False is the
    'did not converge' exit whose gradient _minimize_bfgs recomputes."""
    alpha0 = 0
    if derphi0 != 0:
        alpha1 = min(1.0, 1.01 * 2 * (phi0 - old_phi0) / derphi0)
    else:
        alpha1 = 1.0
    if alpha1 < 0:
        alpha1 = 1.0
    alpha1 = min(alpha1, amax)
    phi_a1, derphi_a1 = yield alpha1
    phi_a0, derphi_a0 = phi0, derphi0
    for i in range(maxiter):
        if alpha1 == 0 or alpha0 > amax:
            return None, None, False
        if phi_a1 > phi0 + c1 * alpha1 * derphi0 or (phi_a1 >= phi_a0 and i > 0):
            a, p, d = yield from _zoom(alpha0, alpha1, phi_a0, phi_a1, derphi_a0, phi0, derphi0,
                                       c1, c2)
            return a, p, a is not None
        if abs(derphi_a1) <= -c2 * derphi0:
            return alpha1, phi_a1, True
        if derphi_a1 >= 0:
            a, p, d = yield from _zoom(alpha1, alpha0, phi_a1, phi_a0, derphi_a1, phi0, derphi0,
                                       c1, c2)
            return a, p, a is not None
        alpha2 = min(2 * alpha1, amax)
        alpha0, alpha1, phi_a0, derphi_a0 = alpha1, alpha2, phi_a1, derphi_a1
        phi_a1, derphi_a1 = yield alpha1
    return alpha1, phi_a1, False


# ------------------------------------------------------------- the minimiser
def bfgs_steps(x0s, hess_inv0=None, gtol=1e-5, eps=_EPSILON, c1=1e-4, c2=0.9, maxiter=None,
               progress=None):
    """scipy.optimize.minimize(method='BFGS', jac=None, options=dict(hess_inv0=
    hess_inv0)) for every row of x0s (B, N), as a generator: it yields
    evaluation requests (idx (K,), X (K, N)) -- rows X of problems idx, N + 1
    consecutive rows per problem: a point and its forward-difference
    neighbours -- is sent their function values, and returns dict(x (B, N),
    fun (B,), nit (B,), status (B,) [scipy's warnflag], success (B,),
    rounds).  `progress`: a dict kept up to date with x (the array of current
    points) and live (problems still searching); the rows of x of the others are
    final, so a driver may hand them on before the slowest problem stops."""
    x = np.array(x0s, dtype=np.float64)
    B, N = x.shape
    if maxiter is None:
        maxiter = N * 200
    eye = np.eye(N, dtype=int)
    H = np.tile(np.asarray(eye if hess_inv0 is None else hess_inv0, dtype=np.float64), (B, 1, 1))
    amin, amax, xtol = 1e-100, 1e100, 1e-14
    ls = _Dcsrch(B, c1, c2, xtol, amin, amax)

    def fg(idx, pts):
        """Function value and gradient at pts (m, N) of problems idx."""
        P, den = fd_points(pts, eps)
        vals = (yield np.repeat(idx, N + 1), P.reshape(-1, N))
        vals = np.asarray(vals, dtype=np.float64).reshape(len(idx), N + 1)
        return vals[:, 0], (vals[:, 1:] - vals[:, :1]) / den

    allb = np.arange(B)
    fval, g = yield from fg(allb, x)
    with np.errstate(all='ignore'):
        old_old = fval + np.sqrt(np.matmul(g[:, None, :], g[:, :, None])[:, 0, 0]) / 2
    nit = np.zeros(B, dtype=np.int64)
    status = np.zeros(B, dtype=np.int64)
    gnorm = np.max(np.abs(g), axis=1)
    pk = np.zeros((B, N))
    derphi0 = np.zeros(B)
    stp = np.zeros(B)
    ls_it = np.zeros(B, dtype=np.int64)
    fallback = {}          # problem -> wolfe2 coroutine
    is_fb = np.zeros(B, dtype=bool)
    rounds = 1

    def begin_search(s):
        """Direction and first step of a new line search for problems s; returns
        the problems whose search could not start (-> fallback search)."""
        if len(s) == 0:
            return s
        with np.errstate(all='ignore'):
            pk[s] = -np.matmul(H[s], g[s][:, :, None])[:, :, 0]
            d0 = np.matmul(g[s][:, None, :], pk[s][:, :, None])[:, 0, 0]
            derphi0[s] = d0
            a1 = _pymin(np.full(len(s), 1.0), 1.01 * 2 * (fval[s] - old_old[s]) / d0)
            a1 = np.where(a1 < 0, 1.0, a1)
            a1 = np.where(d0 != 0, a1, 1.0)
        stp[s] = a1
        ls_it[s] = 1
        task = ls.start(s, a1, fval[s], d0)
        return s[task == _ERROR]

    def to_fallback(s):
        """Start line_search_wolfe2 for problems s; returns those still searching."""
        live = []
        for b in s:
            co = _wolfe2(fval[b], old_old[b], derphi0[b], c1, c2, amax)
            try:
                stp[b] = next(co)
                fallback[int(b)] = co
                is_fb[b] = True
                live.append(b)
            except StopIteration:          # cannot happen before the first value
                status[b] = 2
        return np.array(live, dtype=np.int64)

    # problems whose loop condition holds at the start
    searching = allb[(gnorm > gtol) & (nit < maxiter)]
    failed = begin_search(searching)
    if len(failed):
        searching = np.concatenate([np.setdiff1d(searching, failed), to_fallback(failed)])
    while len(searching):
        s = np.sort(searching)
        if progress is not None:
            progress.update(x=x, live=s)
        pts = x[s] + stp[s][:, None] * pk[s]
        f1, g1 = yield from fg(s, pts)
        rounds += 1
        with np.errstate(all='ignore'):
            d1 = np.matmul(g1[:, None, :], pk[s][:, :, None])[:, 0, 0]
        in_fb = is_fb[s]
        done_ix, done_alpha, done_f, done_g = [], [], [], []
        again = []
        # --- MINPACK search
        m = np.nonzero(~in_fb)[0]
        if len(m):
            sm = s[m]
            nstp, task = ls.iterate(sm, stp[sm], f1[m], d1[m])
            ls_it[sm] += 1
            ok = task == _CONV
            cont = (task == _FG) & np.isfinite(nstp) & (ls_it[sm] < 100)
            fail = ~ok & ~cont
            stp[sm[cont]] = nstp[cont]
            again.append(sm[cont])
            done_ix.append(sm[ok])
            done_alpha.append(stp[sm[ok]])
            done_f.append(f1[m[ok]])
            done_g.append(g1[m[ok]])
            if fail.any():
                again.append(to_fallback(sm[fail]))
        # --- fallback search (per problem)
        for j in np.nonzero(in_fb)[0]:
            b = int(s[j])
            try:
                stp[b] = fallback[b].send((f1[j], d1[j]))
                again.append(np.array([b], dtype=np.int64))
            except StopIteration as stop:
                del fallback[b]
                is_fb[b] = False
                alpha, phi_star, _ = stop.value
                if alpha is None:
                    status[b] = 2          # _LineSearchError -> warnflag 2
                    continue
                # the last point evaluated is the accepted one (see _wolfe2)
                done_ix.append(np.array([b], dtype=np.int64))
                done_alpha.append(np.array([alpha]))
                done_f.append(np.array([phi_star]))
                done_g.append(g1[j][None])
        # --- quasi-Newton update of the problems whose search ended
        u = np.concatenate(done_ix) if done_ix else np.zeros(0, dtype=np.int64)
        if len(u):
            alpha = np.concatenate(done_alpha)
            fnew = np.concatenate(done_f)
            gnew = np.concatenate(done_g)
            with np.errstate(all='ignore'):
                sk = alpha[:, None] * pk[u]
                x[u] = x[u] + sk
                yk = gnew - g[u]
                g[u] = gnew
                old_old[u] = fval[u]
                fval[u] = fnew
                nit[u] += 1
                gn = np.max(np.abs(gnew), axis=1)
                gnorm[u] = gn
                stop = gn <= gtol
                # xrtol = 0: alpha * |pk| <= 0
                stop |= alpha * np.max(np.abs(pk[u]), axis=1) <= 0
                notfin = ~stop & ~np.isfinite(fnew)
                status[u[notfin]] = 2
                stop |= notfin
                rinv = np.matmul(yk[:, None, :], sk[:, :, None])[:, 0, 0]
                rho = np.where(rinv == 0., 1000.0, 1. / rinv)
                A1 = eye - sk[:, :, None] * yk[:, None, :] * rho[:, None, None]
                A2 = eye - yk[:, :, None] * sk[:, None, :] * rho[:, None, None]
                Hn = np.matmul(A1, np.matmul(H[u], A2)) + \
                    rho[:, None, None] * sk[:, :, None] * sk[:, None, :]
            upd = ~stop
            H[u[upd]] = Hn[upd]
            nxt = u[upd & (gn > gtol) & (nit[u] < maxiter)]
            failed = begin_search(nxt)
            if len(failed):
                nxt = np.concatenate([np.setdiff1d(nxt, failed), to_fallback(failed)])
            again.append(nxt)
        searching = np.concatenate(again) if again else np.zeros(0, dtype=np.int64)
    status = np.where(status == 2, 2, np.where(nit >= maxiter, 1, np.where(
        np.isnan(gnorm) | np.isnan(fval) | np.isnan(x).any(axis=1), 3, 0)))
    return dict(x=x, fun=fval, nit=nit, status=status, success=status == 0, rounds=rounds)


class BFGSStepper:
    """The library's host-side lock-step BFGS stepper (csrc/bfgs_host.cpp, rvs_bfgs_*):
    the algorithm of bfgs_steps one problem at a time in C++ -- microseconds per round
    instead of a millisecond of numpy, and steppable without the interpreter
    (LikelihoodEngine.drive_run).  Its matrix products are summed in index order, scipy's
    (and bfgs_steps') go through BLAS, so the two agree to rounding, not bit for bit.
    request() -> (idx, X) or None, feed(values), active(), result()."""

    def __init__(self, x0s, hess_inv0=None, gtol=1e-5, maxiter=None):
        import ctypes
        from . import _cabi, _dev
        self._dev = _dev
        self.L = _cabi.lib()
        x0 = np.ascontiguousarray(x0s, dtype=np.float64)
        self.B, self.N = x0.shape
        h0 = None if hess_inv0 is None else np.ascontiguousarray(hess_inv0, dtype=np.float64)
        h = self.L.rvs_bfgs_create(self.B, self.N, _dev.hptr(x0),
                                   None if h0 is None else _dev.hptr(h0), float(gtol),
                                   int(maxiter or 0))
        if not h:
            raise _cabi.RvsError('rvs_bfgs_create failed')
        self.h = ctypes.c_void_p(h)
        self.cap = max(1, self.B * (self.N + 1))
        self.idx = np.empty(self.cap, dtype=np.int32)
        self.X = np.empty((self.cap, self.N), dtype=np.float64)
        self.n = 0

    def request(self):
        n = self.L.rvs_bfgs_request(self.h, self._dev.hptr(self.idx), self._dev.hptr(self.X),
                                    self.cap)
        assert n <= self.cap
        self.n = n
        return (self.idx[:n], self.X[:n]) if n else None

    def feed(self, f):
        from . import _cabi
        f = np.ascontiguousarray(f, dtype=np.float64)
        _cabi.check(self.L.rvs_bfgs_feed(self.h, self._dev.hptr(f), self.n), 'rvs_bfgs_feed')

    def active(self):
        a = np.empty(self.B, dtype=np.uint8)
        self.L.rvs_bfgs_live(self.h, self._dev.hptr(a))
        return a.astype(bool)

    def result(self):
        import ctypes
        B, N = self.B, self.N
        x, fun = np.empty((B, N)), np.empty(B)
        nit, status = np.empty(B, dtype=np.int64), np.empty(B, dtype=np.int32)
        rounds = ctypes.c_int64(0)
        self.L.rvs_bfgs_result(self.h, self._dev.hptr(x), self._dev.hptr(fun), self._dev.hptr(nit),
                               self._dev.hptr(status), ctypes.byref(rounds))
        return dict(x=x, fun=fun, nit=nit, status=status.astype(np.int64), success=status == 0,
                    rounds=int(rounds.value))

    def close(self):
        if self.h is not None:
            self.L.rvs_bfgs_destroy(self.h)
            self.h = None


def bfgs_native(x0s, hess_inv0=None, gtol=1e-5, maxiter=None):
    """bfgs_steps with the stepping done by BFGSStepper: the same generator protocol."""
    st = BFGSStepper(x0s, hess_inv0, gtol, maxiter)
    try:
        while True:
            req = st.request()
            if req is None:
                break
            st.feed((yield req))
        return st.result()
    finally:
        st.close()


def bfgs_lockstep(fbatch, x0s, hess_inv0=None, native=False, **kw):
    """bfgs_steps (or its native sibling) driven with a blocking objective
    fbatch(idx, X) -> f."""
    gen = (bfgs_native if native else bfgs_steps)(x0s, hess_inv0, **kw)
    try:
        req = next(gen)
        while True:
            req = gen.send(fbatch(*req))
    except StopIteration as stop:
        return stop.value
