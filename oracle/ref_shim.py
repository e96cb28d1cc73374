"""TEST INFRASTRUCTURE ONLY — import shim for the *unmodified* reference at /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
`oracle/make_golden.py` to generate the committed fixtures under tests/golden/ and by
`tests/test_oracle_vs_reference.py` (skipped when the reference is absent) to pin the
restatement in `oracle/met2_oracle.py`.

Nothing is copied from the reference: the modules are imported from where they lie.
The shim provides what the reference's pinned environment (scipy 1.5.2, Python 2-era
names) had and this image (scipy 1.18.1, Python 3.12) lacks:

  1. `scipy.optimize.__nnls` / `scipy.optimize._nnls` modules whose
     `nnls(A, m, n, b, w, zz, index, maxiter)` forwards to the C Lawson-Hanson routine
     `scipy.optimize._slsqplib.nnls(A, b, itmax)` with itmax = 3n when maxiter == -1 and
     returns `(x, rnorm, mode)`; calling `_slsqplib.nnls` directly keeps the reference's
     "no error when itmax is hit" behaviour (intravoxel_algorithms/algorithms.py:77-81).
  2. `builtins.xrange = range` (flip_angle_algorithms/fa_estimation.py:99).
  3. empty stub modules for packages that are only used for I/O and plotting
     (nibabel, progressbar, matplotlib, skimage) so that
     `motor.motor_recon_met2_real_data` can be imported for `fitting_slice_T2` and
     `create_Laplacian_matrix`.
"""
import builtins
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MET2_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "epg", "epg.py"))


def _install_nnls_shim():
    import numpy as np
    import scipy.optimize
    from scipy.optimize import _slsqplib

    def nnls(A, m, n, b, w, zz, index, maxiter):
        itmax = 3 * n if maxiter == -1 else maxiter
        x, rnorm, info = _slsqplib.nnls(np.ascontiguousarray(A, dtype=np.float64),
                                        np.ascontiguousarray(b, dtype=np.float64), itmax)
        return x, rnorm, (1 if info == 0 else 3)

    for name in ("__nnls", "_nnls"):
        full = "scipy.optimize." + name
        mod = types.ModuleType(full)
        mod.nnls = nnls
        # keep scipy's own private module reachable under its real name for scipy itself
        if name == "_nnls" and full in sys.modules:
            real = sys.modules[full]
            if not hasattr(real, "nnls_shim"):
                real.nnls_shim = nnls
            continue
        sys.modules[full] = mod
        setattr(scipy.optimize, name, mod)


def _install_stub_modules():
    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
        return mod

    stub("nibabel")
    stub("progressbar", progressbar=lambda it, **kw: it)
    mpl = stub("matplotlib", rcParams={})
    mpl.pyplot = stub("matplotlib.pyplot")
    mpl.ticker = stub("matplotlib.ticker")
    sk = stub("skimage")
    sk.restoration = stub("skimage.restoration", estimate_sigma=None, denoise_tv_chambolle=None)


_loaded = {}


def load_reference():
    """Return a dict of the reference's hot-path modules (imported read-only)."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    builtins.xrange = range
    _install_nnls_shim()
    _install_stub_modules()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import epg.epg as epg
        import intravoxel_algorithms.algorithms as algorithms
        import intravoxel_algorithms.bayesian_interpolation as bayes
        import flip_angle_algorithms.fa_estimation as fa
        import motor.motor_recon_met2_real_data as motor
    _loaded.update(epg=epg, algorithms=algorithms, bayes=bayes, fa=fa, motor=motor)
    return _loaded
