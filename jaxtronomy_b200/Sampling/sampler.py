"""Batch boundary of the samplers: jaxtronomy/Sampling/sampler.py:215-251 (prepare_logL_func) and the PSO fitness
call jaxtronomy/Sampling/Samplers/pso.py:51-66.  The sampler state update stays on the host (NumPy), exactly as in
the reference; one `positions[B, n]` matrix goes down and one `logL[B]` vector comes back per iteration."""

import numpy as np

from jaxtronomy_b200 import parallel

__all__ = ["prepare_logL_func", "ParticleSwarmOptimizer", "DeviceParticleSwarmOptimizer"]


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

        pos = torch.as_tensor(np.ascontiguousarray(args, dtype=np.float64)).to("cuda")
        if world_size == 1:
            return likelihood_module.logL_batch_device(pos).cpu().numpy().flatten()
        full = parallel.sharded_logl(likelihood_module.logL_batch_device, pos, world_size, rank)
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
        rnd = torch.empty((3, self.particleCount, self.param_count), dtype=torch.float64, device=self.position.device)
        for it in range(max_iter):
            fitness = self._get_fitness()
            rnd.uniform_(0.0, 1.0, generator=self._gen)
            _lib.pso_step(self.position, self.velocity, self.personal_best, self.personal_fitness, self.global_best,
                          self.global_fitness, fitness, rnd, c1, c2)
            history[it] = self.global_fitness[0]
        return self.global_best.cpu().numpy(), history.cpu().numpy().tolist()
