import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_voxels():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "voxels.npz")))


@pytest.fixture(scope="session")
def golden_dictionary():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "dictionary.npz")))


@pytest.fixture(scope="session")
def golden_config2():
    """20 480 voxels of the config-2 phantom fitted by the unmodified reference (oracle/make_golden_config2.py)."""
    import numpy as np
    g = dict(np.load(os.path.join(GOLDEN, "config2_subset.npz")))
    sup = np.unpackbits(g["support"], axis=1)[:, :60].astype(bool)
    f = np.zeros(sup.shape)
    f[sup] = g["f_nz"]
    g["f"] = f
    return g


@pytest.fixture(scope="session")
def golden_methods(golden_config2):
    """2 048 config-2 voxels fitted by the unmodified reference with the other methods (oracle/make_golden_methods.py)."""
    import numpy as np
    g = dict(np.load(os.path.join(GOLDEN, "methods_subset.npz")))
    g["sig"] = np.ascontiguousarray(golden_config2["sig"][::int(g["stride"])])

    def spectrum(key, npc):
        sup = np.unpackbits(g[key + "_support"], axis=1)[:, :npc].astype(bool)
        f = np.zeros(sup.shape)
        f[sup] = g[key + "_fnz"]
        return f
    g["spectrum"] = spectrum
    return g


@pytest.fixture(scope="session")
def golden_methods_invt2(golden_config2):
    """The same 2 048 voxels fitted by the unmodified reference with X2 / L_curve / BayesReg and reg_matrix InvT2, plus
    BayesReg-InvT2 on the signals perturbed by 1e-13 (oracle/make_golden_methods_invt2.py)."""
    import numpy as np
    g = dict(np.load(os.path.join(GOLDEN, "methods_invt2_subset.npz")))
    g["sig"] = np.ascontiguousarray(golden_config2["sig"][::int(g["stride"])])
    g["spectrum"] = lambda key: _unpack(g, key, 60)
    return g


def _unpack(g, key, npc):
    import numpy as np
    sup = np.unpackbits(g[key + "_support"], axis=1)[:, :npc].astype(bool)
    f = np.zeros(sup.shape)
    f[sup] = g[key + "_fnz"]
    return f


@pytest.fixture(scope="session")
def golden_plain_wide(golden_config2):
    """Plain NNLS of the 20 480 config-2 voxels on the 60-, 96- and 100-bin grids + the 96-bin spline FA search, all by the
    unmodified reference (oracle/make_golden_r2.py)."""
    import numpy as np
    g = dict(np.load(os.path.join(GOLDEN, "plain_nnls_wide.npz")))
    return dict(sig=golden_config2["sig"], fa_idx=golden_config2["fa_idx"].astype(np.int32),
                f60=_unpack(g, "nnls60", 60), f96=_unpack(g, "nnls96", 96), f100=_unpack(g, "nnls100", 100),
                fa96_idx=g["fa96_idx"].astype(np.int32), fa96_km=g["fa96_km"])


@pytest.fixture(scope="session")
def golden_config4():
    """2 048 voxels at BASELINE.json configs[3] sizes (nTE 48, 100 bins, brute-force FA, BayesReg + InvT2) fitted by the
    unmodified reference, plus the same fit of the 1e-13-perturbed signals (oracle/make_golden_r2.py)."""
    import numpy as np
    g = dict(np.load(os.path.join(GOLDEN, "config4_subset.npz")))
    for key in ("nnls", "bayes", "bayesp"):
        g["f_" + key] = _unpack(g, key, 100)
    return g
