"""Batch boundary of the samplers: jaxtronomy/Sampling/sampler.py:215-251 (prepare_logL_func) and the PSO fitness
call jaxtronomy/Sampling/Samplers/pso.py:51-66.  The sampler state update stays on the host (NumPy), exactly as in
the reference; one `positions[B, n]` matrix goes down and one `logL[B]` vector comes back per iteration."""

import numpy as np

from jaxtronomy_b200 import parallel

__all__ = ["Sampler", "prepare_logL_func", "ParticleSwarmOptimizer", "DeviceParticleSwarmOptimizer", "DeviceEnsembleSampler",
           "mcmc_emcee", "sample_ball_truncated"]


def prepare_logL_func(backend, likelihood_module, threadCount=1, world_size=1, rank=0):
    """-> ``new_logL_func(positions[B, n]) -> np.ndarray[B]``.

    backend "gpu": one lsim_logl launch sequence per call (the reference: jit(vmap(logL_func))).
    backend "cpu": the reference's pmap-over-host-devices path has no counterpart here -- there is no CPU path.
    With world_size > 1 (torch.distributed initialised) every rank evaluates its slice of the batch and the
    per-particle likelihoods are all-gathered, so every rank returns the full vector."""
    if backend == "cpu":
        raise ValueError("jaxtronomy_b200 has no CPU backend; use backend='gpu'")
    if backend != "gpu":
        raise ValueError("backend must be either 'cpu' or 'gpu'")

    def new_logL_func(args):
        import torch

        args = np.ascontiguousarray(args, dtype=np.float64)
        if world_size == 1:
            pos = torch.as_tensor(args).to("cuda")
            return likelihood_module.logL_batch_device(pos).cpu().numpy().flatten()
        # only this rank's rows cross PCIe; the all-gather returns every particle's logL to every rank
        B = args.shape[0]
        lo, hi = parallel.shard_bounds(B, world_size, rank)
        full = torch.empty(B, dtype=torch.float64, device="cuda")
        if hi > lo:
            full[lo:hi] = likelihood_module.logL_batch_device(torch.as_tensor(args[lo:hi]).to("cuda"))
        parallel.all_gather_logl(full, world_size, rank, B)
        return full.cpu().numpy().flatten()

    return new_logL_func


class ParticleSwarmOptimizer(object):
    """Host-side PSO (standard constricted update) whose only contact with the likelihood is the batched fitness
    call -- the structure of lenstronomy's PSO that the reference subclasses (Samplers/pso.py:8-66)."""

    def __init__(self, func, low, high, particle_count=25, seed=None):
        self.logL_func = func
        self.low, self.high = np.asarray(low, dtype=float), np.asarray(high, dtype=float)
        self.particleCount = particle_count
        self.param_count = len(self.low)
        self._rng = np.random.default_rng(seed)
        self.position = self._rng.uniform(self.low, self.high, (particle_count, self.param_count))
        self.velocity = np.zeros_like(self.position)
        self.personal_best = self.position.copy()
        self.personal_fitness = np.full(particle_count, -np.inf)
        self.global_best, self.global_fitness = None, -np.inf

    def set_global_best(self, position, velocity=None, fitness=None):
        self.global_best = np.asarray(position, dtype=float).copy()
        self.global_fitness = -np.inf if fitness is None else fitness

    def _get_fitness(self):
        return np.asarray(self.logL_func(self.position))

    def optimize(self, max_iter=100, c1=1.193, c2=1.193, verbose=False):
        """-> (best position, [fitness history], [position history])"""
        history, positions = [], []
        for it in range(max_iter):
            fitness = self._get_fitness()
            better = fitness > self.personal_fitness
            self.personal_best[better] = self.position[better]
            self.personal_fitness[better] = fitness[better]
            k = int(np.argmax(self.personal_fitness))
            if self.personal_fitness[k] > self.global_fitness:
                self.global_fitness = float(self.personal_fitness[k])
                self.global_best = self.personal_best[k].copy()
            history.append(self.global_fitness)
            positions.append(self.global_best.copy())
            w = 0.5 + self._rng.uniform(0, 1, self.position.shape) / 2  # per particle and dimension
            r1 = self._rng.uniform(0, 1, self.position.shape)
            r2 = self._rng.uniform(0, 1, self.position.shape)
            self.velocity = (w * self.velocity + c1 * r1 * (self.personal_best - self.position)
                             + c2 * r2 * (self.global_best - self.position))
            self.position = self.position + self.velocity
            if verbose and it % 10 == 0:
                print("iteration %d: best logL %.4f" % (it, self.global_fitness))
        return self.global_best, history, positions


class DeviceParticleSwarmOptimizer(object):
    """GPU-resident particle swarm: positions, velocities and bests live on the device, the
    fitness is ``likelihood_module.logL_batch_device`` on the position matrix itself (the plan's column map is
    Param.args2kwargs), and the swarm update is ``lsim_pso_step``.  Per iteration nothing crosses PCIe; the history
    of the best fitness is read back once at the end.  Same update rule and constants as ParticleSwarmOptimizer
    above (the host-side class that mirrors the reference's interface); the uniform numbers come from torch's
    device generator.

    With world_size > 1 every rank holds the full swarm, evaluates its slice of the particles and all-gathers the
    fitness (parallel.sharded_logl); the update is replicated -- identical on every rank because the random
    numbers are drawn from identically seeded generators."""

    def __init__(self, likelihood_module, low, high, particle_count=25, seed=0, world_size=1, rank=0):
        import torch

        self._like = likelihood_module
        self._world, self._rank = int(world_size), int(rank)
        dev = torch.device("cuda", torch.cuda.current_device())
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(int(seed))
        low = torch.as_tensor(np.asarray(low, dtype=np.float64), device=dev)
        high = torch.as_tensor(np.asarray(high, dtype=np.float64), device=dev)
        self.particleCount, self.param_count = int(particle_count), int(low.numel())
        u = torch.rand((self.particleCount, self.param_count), dtype=torch.float64, device=dev, generator=self._gen)
        self.position = (low + u * (high - low)).contiguous()
        self.velocity = torch.zeros_like(self.position)
        self.personal_best = self.position.clone()
        self.personal_fitness = torch.full((self.particleCount,), -float("inf"), dtype=torch.float64, device=dev)
        self.global_best = self.position[0].clone()
        self.global_fitness = torch.full((1,), -float("inf"), dtype=torch.float64, device=dev)

    def set_global_best(self, position, velocity=None, fitness=None):
        import torch

        self.global_best.copy_(torch.as_tensor(np.asarray(position, dtype=np.float64)))
        self.global_fitness.fill_(-float("inf") if fitness is None else float(fitness))

    def _get_fitness(self):
        if self._world == 1:
            return self._like.logL_batch_device(self.position).contiguous()
        return parallel.sharded_logl(self._like.logL_batch_device, self.position, self._world, self._rank)

    def optimize(self, max_iter=100, c1=1.193, c2=1.193, verbose=False):
        """-> (best position [n] NumPy, [best fitness per iteration])"""
        import torch

        from jaxtronomy_b200 import _lib

        history = torch.empty(max_iter, dtype=torch.float64, device=self.position.device)
        best_pos = torch.empty((max_iter, self.param_count), dtype=torch.float64, device=self.position.device)
        rnd = torch.empty((3, self.particleCount, self.param_count), dtype=torch.float64, device=self.position.device)
        for it in range(max_iter):
            fitness = self._get_fitness()
            rnd.uniform_(0.0, 1.0, generator=self._gen)
            _lib.pso_step(self.position, self.velocity, self.personal_best, self.personal_fitness, self.global_best,
                          self.global_fitness, fitness, rnd, c1, c2)
            history[it] = self.global_fitness[0]
            best_pos[it] = self.global_best
        self.best_position_history = best_pos  # [max_iter, n] device tensor: the global best after every iteration
        return self.global_best.cpu().numpy(), history.cpu().numpy().tolist()


class DeviceEnsembleSampler(object):
    """Affine-invariant ensemble sampler with the stretch move (Goodman & Weare 2010; a = 2, the walkers split at
    random into two halves that are updated in turn against each other): what
    ``emcee.EnsembleSampler(n_walkers, n_param, logL_func, vectorize=True).run_mcmc`` does behind
    ``Sampler.mcmc_emcee`` (jaxtronomy/Sampling/sampler.py:112-212), which hands emcee the batched ``logL_func``.

    Here the walkers, the proposals and the accept / reject step stay on the device: a step is two batched likelihood
    evaluations of n_walkers / 2 proposals each (``LikelihoodModule.logL_batch_device`` on the proposal matrix itself)
    and a handful of element-wise tensor operations; the chain is read back once at the end.  Per half-ensemble S with
    complement C:  z = ((a-1) u + 1)^2 / a,  y_k = c_j + z (x_k - c_j) with j uniform in C,  accept if
    ln u' < (n-1) ln z + logp(y_k) - logp(x_k).

    emcee is not installable here (parity unpinned): same move, same acceptance rule, torch's generator instead of
    NumPy's.  ``logp_fn`` may be any callable on a [B, n] tensor (the CPU tests sample a Gaussian with it).
    """

    def __init__(self, logp_fn, n_walkers, n_param, seed=0, a=2.0, device=None):
        import torch

        if n_walkers < 2 or n_walkers % 2:
            raise ValueError("the stretch move needs an even number of walkers (>= 2)")
        self._logp = getattr(logp_fn, "logL_batch_device", logp_fn)
        self.n_walkers, self.n_param, self.a = int(n_walkers), int(n_param), float(a)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(seed))
        self.chain = None
        self.log_prob = None
        self.n_accepted = None

    def run_mcmc(self, initpos, n_steps):
        """Advance the ensemble by n_steps from ``initpos`` [n_walkers, n_param]; stores chain [n_steps, n_walkers, n_param]
        and log_prob [n_steps, n_walkers] (device tensors)."""
        import torch

        dev, W, n = self.device, self.n_walkers, self.n_param
        x = torch.as_tensor(np.asarray(initpos, dtype=np.float64) if not isinstance(initpos, torch.Tensor) else initpos,
                            dtype=torch.float64, device=dev).clone().contiguous()
        if x.shape != (W, n):
            raise ValueError("initpos must be [%d, %d]" % (W, n))
        lp = self._logp(x).clone()
        if torch.isnan(lp).any():
            raise ValueError("The initial log-probability contains NaN")
        chain = torch.empty((n_steps, W, n), dtype=torch.float64, device=dev)
        logp = torch.empty((n_steps, W), dtype=torch.float64, device=dev)
        accepted = torch.zeros(W, dtype=torch.int64, device=dev)
        half = W // 2
        for it in range(n_steps):
            perm = torch.randperm(W, device=dev, generator=self._gen)
            for first in (True, False):
                s_idx = perm[:half] if first else perm[half:]
                c_idx = perm[half:] if first else perm[:half]
                s, c = x[s_idx], x[c_idx]
                u = torch.rand(half, dtype=torch.float64, device=dev, generator=self._gen)
                zz = ((self.a - 1.0) * u + 1.0) ** 2 / self.a
                rint = torch.randint(half, (half,), device=dev, generator=self._gen)
                cj = c[rint]
                q = cj - (cj - s) * zz[:, None]
                new_lp = self._logp(q.contiguous())
                lnpdiff = (n - 1.0) * torch.log(zz) + new_lp - lp[s_idx]
                acc = torch.log(torch.rand(half, dtype=torch.float64, device=dev, generator=self._gen)) < lnpdiff
                x[s_idx] = torch.where(acc[:, None], q, s)
                lp[s_idx] = torch.where(acc, new_lp, lp[s_idx])
                accepted[s_idx] += acc.to(torch.int64)
            chain[it] = x
            logp[it] = lp
        self.chain, self.log_prob, self.n_accepted = chain, logp, accepted
        return x, lp

    @property
    def acceptance_fraction(self):
        return (self.n_accepted.double() / self.chain.shape[0]).cpu().numpy()

    def get_chain(self, discard=0, thin=1, flat=False):
        c = self.chain[discard::thin]
        return (c.reshape(-1, self.n_param) if flat else c).cpu().numpy()

    def get_log_prob(self, discard=0, thin=1, flat=False):
        lp = self.log_prob[discard::thin]
        return (lp.reshape(-1) if flat else lp).cpu().numpy()


def sample_ball_truncated(mean, sigma, lower_limit, upper_limit, size, seed=None):
    """Start positions of the walkers: independent normals around ``mean`` truncated to the prior box (lenstronomy's
    sampling_util.sample_ball_truncated, used at sampler.py:185-191; restated -- parity unpinned)."""
    from scipy.stats import truncnorm

    mean, sigma = np.asarray(mean, dtype=float), np.asarray(sigma, dtype=float)
    a, b = (np.asarray(lower_limit) - mean) / sigma, (np.asarray(upper_limit) - mean) / sigma
    return truncnorm.rvs(a, b, loc=mean, scale=sigma, size=(size, mean.size), random_state=seed)


def mcmc_emcee(likelihood_module, n_walkers, n_run, n_burn, mean_start, sigma_start, initpos=None, seed=0,
               mpi=False, progress=False, threadCount=1, backend_filename=None, start_from_backend=False):
    """Sampler.mcmc_emcee (jaxtronomy/Sampling/sampler.py:112-212) on the device-resident ensemble sampler:
    -> (flat samples [n_run * n_walkers, n_param], their log-probabilities), burn-in discarded."""
    if mpi:
        raise ValueError("mpi must be False in JAXtronomy, since parallelization is done through JAX")
    if start_from_backend:
        raise ValueError("start_from_backend must be False in JAXtronomy")
    if backend_filename is not None:
        raise ValueError("backend_filename not supported in JAXtronomy")
    if n_walkers % 2:
        n_walkers += 1
    lower, upper = likelihood_module.param.param_limits()
    num_param = likelihood_module.param.num_param()[0]
    if initpos is None:
        initpos = sample_ball_truncated(mean_start, sigma_start, lower, upper, size=n_walkers, seed=seed)
    sampler = DeviceEnsembleSampler(likelihood_module, n_walkers, num_param, seed=seed)
    sampler.run_mcmc(initpos, n_burn + n_run)
    mcmc_emcee.last_sampler = sampler
    return sampler.get_chain(discard=n_burn, flat=True), sampler.get_log_prob(discard=n_burn, flat=True)


class Sampler(object):
    """The sampler front end of the reference (jaxtronomy/Sampling/sampler.py:16-212, a lenstronomy ``Sampler`` whose
    ``pso`` and ``mcmc_emcee`` are re-routed to the batched likelihood): same constructor and method signatures, with
    the swarm / the walkers resident on the device (DeviceParticleSwarmOptimizer, DeviceEnsembleSampler)."""

    def __init__(self, likelihoodModule):
        self.chain = likelihoodModule
        self.lower_limit, self.upper_limit = self.chain.param.param_limits()

    def pso(self, n_particles, n_iterations, lower_start=None, upper_start=None, threadCount=1, init_pos=None, mpi=False,
            print_key="PSO", verbose=False, seed=0):
        """-> (best-fit argument vector, [logL of the global best per iteration, its position per iteration, []])
        (sampler.py:20-110; the reference's third list holds velocities, which stay on the device here)."""
        if mpi:
            raise ValueError("mpi must be False in JAXtronomy since parallelization is done through JAX")
        if lower_start is None or upper_start is None:
            lower_start, upper_start = np.array(self.lower_limit), np.array(self.upper_limit)
        else:
            lower_start = np.maximum(lower_start, self.lower_limit)
            upper_start = np.minimum(upper_start, self.upper_limit)
        pso = DeviceParticleSwarmOptimizer(self.chain, lower_start, upper_start, n_particles, seed=seed)
        if init_pos is None:
            init_pos = (upper_start - lower_start) / 2 + lower_start
        pso.set_global_best(init_pos, [0] * len(init_pos), self.chain.logL(init_pos))
        result, log_likelihood_list = pso.optimize(n_iterations, verbose=verbose)
        pos_list = pso.best_position_history.cpu().numpy().tolist()
        if verbose:
            print(log_likelihood_list[-1], "log likelihood of the best position (%s)" % print_key)
        return result, [log_likelihood_list, pos_list, []]

    def mcmc_emcee(self, n_walkers, n_run, n_burn, mean_start, sigma_start, mpi=False, progress=False, threadCount=1,
                   initpos=None, backend_filename=None, start_from_backend=False, seed=0):
        """-> (flat samples, their log-probabilities)   (sampler.py:112-212)"""
        return mcmc_emcee(self.chain, n_walkers, n_run, n_burn, mean_start, sigma_start, initpos=initpos, seed=seed,
                          mpi=mpi, progress=progress, threadCount=threadCount, backend_filename=backend_filename,
                          start_from_backend=start_from_backend)
