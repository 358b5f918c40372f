"""CPU oracle for the forward image-simulation hot path of JAXtronomy.

TEST INFRASTRUCTURE ONLY.  Nothing in ``jaxtronomy_b200`` (the product) may
import, call or link anything from this package; the only permitted users are
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` -- and there only as the checker / the reported
CPU baseline, never as the thing measured or shipped.

What it is: a plain NumPy/SciPy restatement of the reference algorithm (op
order, clamps, ``nmax`` schedule, sentinels), every function citing the
``/root/reference`` file:line it follows.

How it is pinned: the reference is pure Python on top of jax + lenstronomy,
neither of which is installable here.  ``tests/golden/make_golden.py`` executes
the *unmodified reference source files* from ``/root/reference`` with
``jax.numpy`` substituted by NumPy (``tests/golden/refshim``) and writes their
outputs, for the reference's own test parameters, to ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks this oracle against those fixtures and
against ``scipy.special.hyp2f1`` (the reference's only oracle-independent
numeric pin, test/test_Util/test_hyp2f1_util.py).  The lenstronomy host-side
helpers the reference imports (grids, PixelGrid, PSF, kernel cut) are absent
from ``/root/reference``; they are restated from lenstronomy's published
behaviour in ``oracle/lenstronomy_port.py`` and are *parity-unpinned* (they are
init-time bookkeeping, not per-pixel arithmetic).
"""
