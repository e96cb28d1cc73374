"""Host-side logic: grids and regularisation matrices, phantom generator, masked-voxel gather, slab partition."""
import numpy as np
import pytest

import met2_oracle as O
from multicomponent_t2_toolbox_b200 import grids, pipeline
from multicomponent_t2_toolbox_b200.phantom import epg_signal_batch, make_phantom, tile_volume


@pytest.mark.parametrize("rm", ["I", "L1", "L2", "InvT2"])
def test_reg_matrices_match_oracle(rm):
    T2s = grids.t2_grid(60)
    L = grids.reg_matrix(rm, T2s)
    assert np.array_equal(L, O._grids("X2", rm, "spline", 40.0, 32, 10.0, 1000.0)["L"])
    K = L.T @ L
    off = np.abs(np.subtract.outer(np.arange(60), np.arange(60))) > 2
    assert not K[off].any()                       # what the 5-band kernels rely on


def test_grids_match_reference_constants():
    T2s = grids.t2_grid(60)
    m, t, c = grids.compartment_masks(T2s, 40.0)
    assert (m.sum(), t.sum(), c.sum()) == (16, 18, 26)
    assert grids.default_npc("T2SPARC") == 96 and grids.default_npc("X2") == 60
    a, k = grids.fa_grids("spline")
    assert len(a) == 273 and len(k) == 15 and a[0] == 90.0 and a[-1] == 180.0
    a, k = grids.fa_grids("brute-force")
    assert len(a) == 91 and k is None
    lam = grids.lambda_grid()
    assert lam[0] == 0.0 and np.isclose(lam[1], 1e-8) and np.isclose(lam[-1], 10.0) and len(lam) == 50
    with pytest.raises(ValueError):
        grids.reg_matrix("L3", T2s)
    with pytest.raises(ValueError):
        grids.fa_grids("newton")


def test_phantom_epg_matches_oracle_epg():
    rad = np.pi / 180.0
    for a in (95.0, 131.7, 180.0):
        for T2 in (10.0, 73.0, 2000.0):
            ref = O.epg_signal(32, 10.0, 1.0 / 1000.0, 1.0 / T2, a * rad, a / 2 * rad)
            got = epg_signal_batch(32, 10.0, [1000.0], [T2], [a])[0]
            assert np.abs(ref - got).max() <= 1e-14


def test_phantom_is_seeded_and_masked():
    p1 = make_phantom((6, 5, 2), seed=3, mask_mode="ellipsoid")
    p2 = make_phantom((6, 5, 2), seed=3, mask_mode="ellipsoid")
    assert np.array_equal(p1["data"], p2["data"])
    assert p1["data"].shape == (6, 5, 2, 32) and not p1["data"][p1["mask"] == 0].any()
    d, m = tile_volume(p1["data"], p1["mask"], (2, 1, 1))
    assert d.shape == (12, 5, 2, 32) and m.shape == (12, 5, 2)


def test_masked_voxel_list_order_and_clamp():
    data = np.arange(2 * 3 * 2 * 4, dtype=float).reshape(2, 3, 2, 4) - 5.0
    mask = np.zeros((2, 3, 2), dtype=np.int64)
    mask[0, 1, 1] = 1
    mask[1, 2, 0] = 1
    flat, sig = pipeline.masked_voxel_list(data, mask)
    assert list(flat) == [3, 10]
    assert np.array_equal(sig[0], np.maximum(data[0, 1, 1], 0.0)) and sig.min() >= 0.0


@pytest.mark.parametrize("V,W", [(0, 4), (1, 8), (10, 3), (552960, 8), (7, 7), (5, 8)])
def test_slab_bounds_partition(V, W):
    bounds = [pipeline.slab_bounds(V, r, W) for r in range(W)]
    covered = []
    for lo, hi in bounds:
        assert 0 <= lo <= hi <= V
        covered.extend(range(lo, hi))
    assert covered == list(range(V))
    sizes = [hi - lo for lo, hi in bounds]
    assert max(sizes) - min(s for s in sizes if s > 0 or V == 0) <= max(sizes)  # contiguous, ceil-sized slabs
    with pytest.raises(ValueError):
        pipeline.slab_bounds(V, W, W)


def test_cyclic_slab_partition():
    from multicomponent_t2_toolbox_b200 import pipeline
    for n, w in [(0, 2), (5, 3), (10000, 4), (552960, 8)]:
        parts = [pipeline.cyclic_slab(n, r, w, chunk=2048) for r in range(w)]
        allidx = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
        assert np.array_equal(np.sort(allidx), np.arange(n))
        if n >= 2048 * w * 4:
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 2048
