"""Development aid: per-parameter table of the device gradient against oracle finite differences at several steps."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tests.golden_io import load
from tests.test_gpu_grad import _base_row, _fd, _oracle_logl, _perturbed
from tests.test_gpu_parity import build_image_model
from tests.test_oracle_golden import build_oracle_image_model

IMSIM = load("imsim.npz")
for name in sys.argv[1:] or ["headline_epl_shear_pixel", "c1_sie_shear_gauss", "linear_sis_3comp"]:
    c = IMSIM[name]
    lin = build_image_model(c, linear=True)
    om = build_oracle_image_model(c)
    names = lin.param_names
    mat = _perturbed(_base_row(c, names), names, np.random.default_rng(5), 3)
    t = torch.as_tensor(mat, device="cuda")
    for linear in (True, False):
        grad = (lin.logL_linear_grad_batch(t)[1] if linear else lin.logL_grad_batch(t)[1]).cpu().numpy()
        b = 0
        row = mat[b]
        print("==", name, "linear" if linear else "joint", "psf", c.meta["kwargs_psf"].get("psf_type"))
        for j, (comp, i, n) in enumerate(names):
            if linear and n == "amp":
                continue
            f = lambda r: _oracle_logl(om, c, names, r, linear)
            fds = [_fd(f, row, j, h) for h in (1e-3, 1e-4, 1e-5)]
            e = np.zeros_like(row)
            e[j] = 1e-6
            fd2 = (f(row + e) - f(row - e)) / 2e-6
            print("%-10s %d %-9s grad %+.10e  fd1e-3 %+.10e fd1e-4 %+.10e fd1e-5 %+.10e fd2_1e-6 %+.10e"
                  % (comp, i, n, grad[b, j], fds[0], fds[1], fds[2], fd2))
