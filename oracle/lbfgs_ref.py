"""NumPy float64 restatement of the L-BFGS defined in csrc/lbfgs.cu (oracle; test infra).

PARITY UNPINNED against the reference: its minimiser is edu.umass.nlp.optimize.LBFGSMinimizer from the
un-vendored umass-nlp-1.0-SNAPSHOT.jar (call site /root/reference/src/Run_NTN.java:119-124, maxIters = 5,
fresh instance per batch job); the jar's source is not in the tree, so this file and lbfgs.cu define the
algorithm: two-loop recursion (history m=6, H0 = s.y/y.y), first step unit-norm steepest descent,
backtracking (Armijo 1e-4, shrink 0.5, 0.1 on the first iteration, <= 20 trials), relative-change stop."""
import numpy as np


def minimize(fn, theta0, max_iters=5, history=6, tol=1e-4):
    """fn(theta) -> (cost, grad).  Returns (theta, f, iters, evals, converged)."""
    theta = np.array(theta0, np.float64)
    f, g = fn(theta)
    evals = 1
    S, Y, rho = [], [], []
    converged = False
    it = 0
    for it in range(max_iters):
        q = g.copy()
        alpha = []
        for s, y, r in zip(reversed(S), reversed(Y), reversed(rho)):
            a = r * (s @ q)
            alpha.append(a)
            q -= a * y
        if S:
            q *= 1.0 / (rho[-1] * (Y[-1] @ Y[-1]))
        for (s, y, r), a in zip(zip(S, Y, rho), reversed(alpha)):
            q += (a - r * (y @ q)) * s
        gd = -(g @ q)
        t, shrink = 1.0, 0.5
        if not S:
            t, shrink = 1.0 / np.sqrt(-gd), 0.1
        if not gd < 0:
            q = g.copy()
            gd = -(g @ g)
            S, Y, rho = [], [], []
            if not gd < 0:
                converged = True
                break
            t, shrink = 1.0 / np.sqrt(-gd), 0.1
        ok = False
        for _ in range(20):
            trial = theta - t * q
            fn_, gn = fn(trial)
            evals += 1
            if fn_ <= f + 1e-4 * t * gd:
                ok = True
                break
            t *= shrink
        if not ok:
            break
        s, y = trial - theta, gn - g
        ys = y @ s
        if ys > 0 and np.isfinite(1.0 / ys):
            S.append(s); Y.append(y); rho.append(1.0 / ys)
            if len(S) > history:
                S.pop(0); Y.pop(0); rho.pop(0)
        theta, g = trial, gn
        fold, f = f, fn_
        if abs(fold - f) <= tol * (abs(fold) + abs(f) + 1e-10) / 2.0:
            converged = True
            it += 1
            break
    else:
        it = max_iters
    return theta, f, it, evals, converged
