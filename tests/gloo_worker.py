"""Worker of test_all_gather_logl_gloo_world2: python gloo_worker.py RANK WORLD N"""
import sys

import torch
import torch.distributed as dist

from jaxtronomy_b200 import parallel

rank, world, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dist.init_process_group("gloo", rank=rank, world_size=world)
params = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
fn = lambda p: (p**2).sum(dim=1) - 4.0  # noqa: E731
full = parallel.sharded_logl(fn, params, world, rank)
assert torch.equal(full, fn(params)), (rank, full)
lo, hi = parallel.shard_bounds(n, world, rank)
assert hi - lo in (n // world, n // world + 1)
vg = lambda p: ((p**2).sum(dim=1) - 4.0, 2.0 * p)  # noqa: E731
l, g = parallel.sharded_logl_grad(vg, params, world, rank)
assert torch.equal(l, fn(params)) and torch.equal(g, 2.0 * params), (rank, l, g)
dist.destroy_process_group()
print("OK", rank)
