"""ORACLE (test infrastructure): NumPy restatement of one particle-swarm iteration.

Follows the loop body of lenstronomy's ``ParticleSwarmOptimizer._sample`` (lenstronomy/Sampling/Samplers/pso.py,
lenstronomy >= 1.13.6 -- NOT under /root/reference; jaxtronomy/Sampling/Samplers/pso.py:11-66 subclasses it and only
replaces ``_get_fitness`` by one vectorised likelihood call).  PARITY UNPINNED: lenstronomy cannot be imported here,
the restatement is from knowledge of that file:

    for particle in swarm:
        if global_best.fitness < particle.fitness: global_best = particle.copy()
        if particle.fitness > particle.personal_best.fitness: particle.update_personal_best()
    for particle in swarm:
        w = 0.5 + uniform(0, 1, size=param_count) / 2
        part_vel = w * particle.velocity
        cog_vel = c1 * uniform(0, 1, size=param_count) * (particle.personal_best.position - particle.position)
        soc_vel = c2 * uniform(0, 1, size=param_count) * (global_best.position - particle.position)
        particle.velocity = part_vel + cog_vel + soc_vel
        particle.position = particle.position + particle.velocity

The random numbers are an input ([3, B, P]: w-draw, cognitive draw, social draw) so that the CUDA kernel can be
compared bit for bit.
"""

import numpy as np


def pso_step(position, velocity, pbest_pos, pbest_fit, gbest_pos, gbest_fit, fitness, rnd, c1=1.193, c2=1.193):
    """All arrays are updated in place (gbest_fit is a 1-element array), like the particle objects of the reference."""
    B, P = position.shape
    for b in range(B):  # sequential scan: the first particle with the largest fitness wins
        if gbest_fit[0] < fitness[b]:
            gbest_fit[0] = fitness[b]
            gbest_pos[:] = position[b]
        if fitness[b] > pbest_fit[b]:
            pbest_fit[b] = fitness[b]
            pbest_pos[b] = position[b]
    w = 0.5 + rnd[0] / 2
    part_vel = w * velocity
    cog_vel = c1 * rnd[1] * (pbest_pos - position)
    soc_vel = c2 * rnd[2] * (gbest_pos[None, :] - position)
    velocity[:] = part_vel + cog_vel + soc_vel
    position[:] = position + velocity
