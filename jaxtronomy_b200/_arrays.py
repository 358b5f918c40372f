"""Input/output conventions of the façade classes.

The reference takes and returns jax arrays.  Here NumPy in -> NumPy out (blocking, one D2H copy), torch CUDA
tensor in -> torch CUDA tensor out (asynchronous on the current stream).  A parameter given as a ``[B]`` array
instead of a scalar makes the call batched: outputs then carry a leading ``B`` axis -- the equivalent of
``jax.vmap`` over that parameter (Sampling/sampler.py:241-246).
"""

import numpy as np


def is_torch(v):
    return type(v).__module__.startswith("torch")


def flatten_coords(x, y):
    """-> (x_flat, y_flat, shape, wants_torch).  x, y must share their shape."""
    if is_torch(x) or is_torch(y):
        import torch

        xt = x if is_torch(x) else torch.as_tensor(np.asarray(x, dtype=np.float64))
        yt = y if is_torch(y) else torch.as_tensor(np.asarray(y, dtype=np.float64))
        if xt.shape != yt.shape:
            raise ValueError("x and y must have the same shape")
        return xt.reshape(-1), yt.reshape(-1), tuple(xt.shape), True
    xa = np.asarray(x, dtype=np.float64)
    ya = np.asarray(y, dtype=np.float64)
    if xa.shape != ya.shape:
        raise ValueError("x and y must have the same shape")
    return xa.reshape(-1), ya.reshape(-1), xa.shape, False


def finish(t, shape, batched, want_torch):
    """device tensor [B, *lead, N] -> user array: N reshaped to `shape`, B dropped unless batched."""
    lead = tuple(t.shape[1:-1])
    out = t.reshape((t.shape[0],) + lead + tuple(shape))
    if not batched:
        out = out[0]
    if want_torch:
        return out
    return out.cpu().numpy()
