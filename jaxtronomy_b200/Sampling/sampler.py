"""Batch boundary of the samplers: jaxtronomy/Sampling/sampler.py:215-251 (prepare_logL_func) and the PSO fitness
call jaxtronomy/Sampling/Samplers/pso.py:51-66.  The sampler state update stays on the host (NumPy), exactly as in
the reference; one `positions[B, n]` matrix goes down and one `logL[B]` vector comes back per iteration."""

import numpy as np

from jaxtronomy_b200 import parallel

__all__ = ["prepare_logL_func", "ParticleSwarmOptimizer"]


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
            w = 0.5 + self._rng.uniform(0, 1, (self.particleCount, 1)) / 2
            r1 = self._rng.uniform(0, 1, self.position.shape)
            r2 = self._rng.uniform(0, 1, self.position.shape)
            self.velocity = (w * self.velocity + c1 * r1 * (self.personal_best - self.position)
                             + c2 * r2 * (self.global_best - self.position))
            self.position = self.position + self.velocity
            if verbose and it % 10 == 0:
                print("iteration %d: best logL %.4f" % (it, self.global_fitness))
        return self.global_best, history, positions
