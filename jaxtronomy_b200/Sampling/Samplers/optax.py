"""Gradient-descent sampler with the interface of jaxtronomy/Sampling/Samplers/optax.py:12-155 (OptaxMinimizer).

The reference minimises ``-logL`` with ``optax.lbfgs`` in the unconstrained space of a numpyro
``TruncatedNormal(mean, sigma, low, high)``, one chain after the other, each step one ``jax.value_and_grad``.  Here the
value and the gradient of ALL chains come from one ``lsim_logl_grad`` call per trial point (batched over the chains),
and the few-hundred-byte L-BFGS state stays on the host in NumPy:

* unconstrained <-> constrained: ``x = low + (high - low) sigmoid(u)`` -- numpyro's ``biject_to(interval(low, high))``
  (what ``constrain_fn`` / ``unconstrain_fn`` apply to a TruncatedNormal site, optax.py:56-61,80-85,134-141); the loss
  is ``-logL(x(u))`` without a Jacobian term, as in the reference (``_loss``, optax.py:126-143);
* stop rule: ``value - new_value < tol`` three iterations in a row, or ``maxiter`` (optax.py:104-113);
* the best chain's constrained parameters are returned (optax.py:72-87).

Not reproduced bit for bit (neither optax nor numpyro is installable here -- parity unpinned): the random initial draw
(NumPy generator + inverse-CDF truncated normal instead of jax's PRNG) and the line search (Armijo backtracking on the
L-BFGS direction instead of optax's zoom line search).  Both walk to the same local minimum of the same loss.
"""

import numpy as np

__all__ = ["OptaxMinimizer"]


def _sigmoid(u):
    return 0.5 * (1.0 + np.tanh(0.5 * u))


class OptaxMinimizer:
    """Gradient descent using L-BFGS on the batched device gradient."""

    def __init__(self, logL_func, args_mean, args_sigma, args_lower, args_upper, maxiter=500, history=10):
        """
        :param logL_func: a LikelihoodModule (its ``logL_and_grad_batch`` is used) or any callable
            ``positions[B, n] -> (logL[B], dlogL/dpositions[B, n])``
        :param args_mean, args_sigma, args_lower, args_upper: truncated-normal description of the start points
            (Param.kwargs2args order), as in the reference
        :param maxiter: maximum number of L-BFGS iterations per chain
        """
        self._vg = getattr(logL_func, "logL_and_grad_batch", logL_func)
        self.mean = np.asarray(args_mean, dtype=float)
        self.sigma = np.asarray(args_sigma, dtype=float)
        self.low = np.asarray(args_lower, dtype=float)
        self.high = np.asarray(args_upper, dtype=float)
        if not np.all(self.high > self.low):
            raise ValueError("args_upper must be larger than args_lower for every parameter")
        self.maxiter = int(maxiter)
        self.history = int(history)
        self.num_evaluations = 0

    # -- constrained <-> unconstrained -------------------------------------------------------------------------------
    def constrain(self, u):
        return self.low + (self.high - self.low) * _sigmoid(u)

    def unconstrain(self, x):
        t = (np.asarray(x, dtype=float) - self.low) / (self.high - self.low)
        t = np.clip(t, 1e-15, 1 - 1e-15)
        return np.log(t) - np.log1p(-t)

    def _loss_and_grad(self, u):
        """[B, n] unconstrained -> (-logL [B], d(-logL)/du [B, n])"""
        s = _sigmoid(u)
        x = self.low + (self.high - self.low) * s
        logl, grad = self._vg(x)
        self.num_evaluations += 1
        logl = np.asarray(logl, dtype=float)
        grad = np.asarray(grad, dtype=float) * ((self.high - self.low) * s * (1.0 - s))
        return -logl, -grad

    def _loss(self, args_unconstrained):
        return self._loss_and_grad(np.atleast_2d(args_unconstrained))[0]

    def _draw_init_params(self, num_chains, rng_seed):
        """Start points from TruncatedNormal(mean, sigma, low, high) by inverse CDF."""
        from scipy.stats import norm

        rng = np.random.default_rng(rng_seed)
        a, b = norm.cdf((self.low - self.mean) / self.sigma), norm.cdf((self.high - self.mean) / self.sigma)
        q = a + (b - a) * rng.uniform(size=(num_chains, self.mean.size))
        x = self.mean + self.sigma * norm.ppf(np.clip(q, 1e-300, 1 - 1e-16))
        span = self.high - self.low
        return np.clip(x, self.low + 1e-9 * span, self.high - 1e-9 * span)

    # -- batched L-BFGS ------------------------------------------------------------------------------------------------
    def _direction(self, g, S, Y):
        """Two-loop recursion per chain; S, Y: lists of [B, n] (most recent last)."""
        q = g.copy()
        alphas = []
        for s, y in zip(reversed(S), reversed(Y)):
            sy = np.sum(s * y, axis=1)
            rho = np.where(sy > 1e-300, 1.0 / np.where(sy > 1e-300, sy, 1.0), 0.0)
            a = rho * np.sum(s * q, axis=1)
            q -= a[:, None] * y
            alphas.append((a, rho))
        if S:
            sy, yy = np.sum(S[-1] * Y[-1], axis=1), np.sum(Y[-1] * Y[-1], axis=1)
            gamma = np.where((sy > 1e-300) & (yy > 1e-300), sy / np.where(yy > 1e-300, yy, 1.0), 1.0)
            q *= gamma[:, None]
        for (a, rho), s, y in zip(reversed(alphas), S, Y):
            b = rho * np.sum(y * q, axis=1)
            q += (a - b)[:, None] * s
        return -q

    def run_chains(self, init_unconstrained, tol):
        """All chains at once.  -> (final unconstrained params [B, n], iterations per chain [B], final loss [B])"""
        u = np.array(init_unconstrained, dtype=float, ndmin=2)
        B = u.shape[0]
        f, g = self._loss_and_grad(u)
        S, Y = [], []
        tol_hit = np.zeros(B, dtype=int)
        iters = np.zeros(B, dtype=int)
        active = np.isfinite(f)
        for _ in range(self.maxiter):
            if not active.any():
                break
            d = self._direction(g, S, Y)
            slope = np.sum(d * g, axis=1)
            bad = ~(slope < 0)  # not a descent direction: steepest descent for that chain
            d[bad] = -g[bad]
            slope[bad] = -np.sum(g[bad] * g[bad], axis=1)
            step = np.ones(B) if S else 1.0 / np.maximum(np.sqrt(np.sum(g * g, axis=1)), 1.0)
            done = ~active
            u_new, f_new, g_new = u.copy(), f.copy(), g.copy()
            for _ls in range(25):  # Armijo backtracking, all unfinished chains in one batched evaluation
                todo = ~done
                if not todo.any():
                    break
                trial = u[todo] + step[todo, None] * d[todo]
                ft, gt = self._loss_and_grad(trial)
                ok = np.isfinite(ft) & (ft <= f[todo] + 1e-4 * step[todo] * slope[todo])
                idx = np.flatnonzero(todo)
                acc = idx[ok]
                u_new[acc], f_new[acc], g_new[acc] = trial[ok], ft[ok], gt[ok]
                done[acc] = True
                step[idx[~ok]] *= 0.5
            moved = done & active
            s_k, y_k = np.where(moved[:, None], u_new - u, 0.0), np.where(moved[:, None], g_new - g, 0.0)
            diff = f - f_new
            stalled = active & ~moved  # line search failed: the chain has converged as far as fp64 allows
            tol_hit = np.where(active & (diff < tol), tol_hit + 1, np.where(active, 0, tol_hit))
            iters += active.astype(int)
            u, f, g = u_new, f_new, g_new
            S.append(s_k)
            Y.append(y_k)
            if len(S) > self.history:
                S.pop(0)
                Y.pop(0)
            active = active & ~stalled & (tol_hit != 3)
        return u, iters, f

    def run_single(self, init_params, tol):
        """One chain (unconstrained start point) -> (final unconstrained params, number of iterations)."""
        u, iters, _ = self.run_chains(np.asarray(init_params, dtype=float)[None, :], tol)
        return u[0], int(iters[0])

    def run(self, num_chains, tol, rng_seed, verbose=False):
        """Runs the gradient descent for ``num_chains`` start points drawn from the truncated normal and returns the
        constrained parameters of the best chain.

        :param tol: a chain stops when |logL[i] - logL[i-1]| < tol three times in a row
        """
        init = self.unconstrain(self._draw_init_params(num_chains, rng_seed))
        u, iters, f = self.run_chains(init, tol)
        best = int(np.nanargmin(np.where(np.isfinite(f), f, np.inf)))
        if verbose:
            for i in range(num_chains):
                print("chain %d: %d iterations, final logL %.6f" % (i + 1, iters[i], -f[i]))
        self.final_logL = -f
        self.num_iterations = iters
        return self.constrain(u[best])
