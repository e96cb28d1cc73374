"""The other entry points of include/met2.h through the SIMT-emulated library (tests/emu): EPG dictionary, Gram / band
tables, FA-stage Gaussian smoothing, NESMA denoiser, segmented means — the kernels' own source on the CPU against the
reference's golden dictionary, SciPy and the oracle restatements (which are pinned to runs of the unmodified reference,
tests/test_oracle_pipeline_golden.py).  The `-m gpu` tests make the same comparisons on the device."""
import numpy as np
import pytest
from scipy import ndimage

import met2_oracle as O
from emu import emu


@pytest.fixture(scope="module", autouse=True)
def _built():
    emu.build()


def test_epg_dictionary_kernel_against_reference_dictionary(golden_dictionary):
    """epg_dictionary_kernel (banded recurrence) vs columns of the dictionary built by the UNMODIFIED reference's dense
    EPG (epg/epg.py:155): 32 echoes x 60 bins and the config-4 size 48 x 100."""
    g = golden_dictionary
    dic, dicT = emu.epg_dictionary(g["alphas"], g["T2s"], g["T1s"], int(g["nte"]), float(g["tau"]), float(g["TR"]))
    ref = np.transpose(g["dic"], (2, 0, 1))
    assert np.max(np.abs(dic - ref) / np.abs(ref).max()) < 1e-13
    assert np.array_equal(dicT, np.transpose(dic, (0, 2, 1)))
    # config-4 size, straight from the reference (oracle/make_golden.py: 100 bins, 48 echoes, tau 8 ms, TR 2 s)
    T2s100 = np.logspace(1.0, np.log10(2000.0), 100)
    d48, _ = emu.epg_dictionary([90.0, 133.0, 180.0], T2s100, 1000.0 * np.ones(100), 48, 8.0, 2000.0)
    ref48 = np.transpose(g["dic48"], (2, 0, 1))
    assert np.max(np.abs(d48 - ref48)) < 1e-13 * np.abs(ref48).max()
    # 180 degrees: EPG reduces to the mono-exponential (1 - exp(-TR/T1)) exp(-TE/T2)  
    d180, _ = emu.epg_dictionary([180.0], g["T2s"], g["T1s"], 32, 10.0, 1000.0)
    te = 10.0 * np.arange(1, 33)
    mono = (1.0 - np.exp(-1000.0 / g["T1s"]))[None, :] * np.exp(-te[:, None] / g["T2s"][None, :])
    assert np.max(np.abs(d180[0] - mono)) < 1e-13


@pytest.mark.parametrize("matrix", ["I", "L1", "L2", "InvT2"])
def test_gram_and_band_tables(matrix, golden_dictionary):
    g = golden_dictionary
    gr = O._grids("X2", matrix, "brute-force", 40.0, 32, 10.0, 1000.0)
    dic = np.ascontiguousarray(np.transpose(g["dic"], (2, 0, 1)))
    G, kband, err = emu.gram_tables(dic, gr["L"])
    _, _, G_np, kband_np = emu.tables(g["dic"], gr["L"])
    assert err == 0
    assert np.max(np.abs(G - G_np)) < 1e-12 * np.abs(G_np).max() and np.array_equal(G, np.transpose(G, (0, 2, 1)))
    assert np.allclose(kband, kband_np, rtol=1e-14, atol=0)
    dense = np.ones((60, 60))                                      # not pentadiagonal: flagged
    assert emu.gram_tables(dic, dense)[2] == 1


def test_gaussian_smooth_bitwise_equal_to_scipy():
    """motor...:336-346: filt.gaussian_filter(data[:, :, :, c], 2.0, 0) per echo."""
    rng = np.random.default_rng(5)
    vol = rng.uniform(0.0, 1000.0, size=(9, 7, 6, 3))
    out = emu.gaussian_smooth(vol, 2.0)
    for c in range(vol.shape[3]):
        assert np.array_equal(out[..., c], ndimage.gaussian_filter(vol[..., c], 2.0, 0))


def test_nesma_kernel_bitwise_equal_to_the_restatement():
    """motor...:305-333 on a volume larger than the 12-voxel window in x, with masked-out, all-zero and isolated voxels."""
    rng = np.random.default_rng(6)
    nx, ny, nz, nt = 15, 6, 5, 32
    base = 1000.0 * np.exp(-np.arange(nt) / 6.0)
    vol = base * (1.0 + 0.01 * rng.standard_normal((nx, ny, nz, nt)))
    vol[3, 2, 1] *= 3.0                                            # no similar neighbour except itself
    mask = (rng.random((nx, ny, nz)) < 0.8).astype(np.int32)
    mask[0, 0, 0] = 1
    vol[0, 0, 0] = 0.0                                             # all-zero voxel inside the mask: 0/0 -> NaN
    mask[5, 3, 2] = 2                                              # mask != 1 -> untouched zeros
    with np.errstate(all="ignore"):
        ref = O.nesma_filter(vol, mask)
    out = emu.nesma_filter(vol, mask)
    assert np.array_equal(out, ref, equal_nan=True)
    assert np.isnan(out[0, 0, 0]).all() and not out[5, 3, 2].any()


def test_segment_means_against_the_restatement(golden_dictionary):
    """Mask-mean / ROI-mean signal and kernel (motor...:377-392, motor_recon_met2_real_data_ROI.py:408-423)."""
    g = golden_dictionary
    rng = np.random.default_rng(7)
    nx, ny, nz, nt = 6, 5, 4, 32
    data = rng.uniform(1.0, 100.0, size=(nx, ny, nz, nt))
    fa = rng.integers(0, 4, size=(nx, ny, nz))
    rois = rng.integers(0, 4, size=(nx, ny, nz))                    # labels 1..3, 0 = background; label 5 is empty
    dic = np.ascontiguousarray(np.transpose(g["dic"], (2, 0, 1)))
    label = rois.reshape(-1).astype(np.int32) - 1                   # segment id, -1 = none
    label = np.where(label == 2, 3, label)                          # leave segment 2 empty
    ms, mk, cnt = emu.segment_means(data.reshape(-1, nt), fa.reshape(-1), label, 4, dic)
    for seg, lab in ((0, 1), (1, 2), (3, 3)):
        s_ref, k_ref, nv = O.segment_mean(data, fa, g["dic"], rois == lab)
        assert cnt[seg] == nv
        assert np.max(np.abs(ms[seg] - s_ref)) < 1e-12 * np.abs(s_ref).max()
        assert np.max(np.abs(mk[seg] - k_ref)) < 1e-12 * np.abs(k_ref).max()
    assert cnt[2] == 0 and np.isnan(ms[2]).all() and np.isnan(mk[2]).all()   # the reference divides by nv = 0


def test_c_abi_argument_checks_and_empty_batches():
    """Error behaviour of the entry points (include/met2.h conventions): negative return + message, nothing launched;
    an empty batch is a successful no-op.  Same host code as libmet2.so."""
    import ctypes
    L = emu.lib()
    L.met2_last_error.restype = ctypes.c_char_p
    n, m = 60, 32
    cfg = emu.T2Cfg(method=2, nTE=m, nT2=n, nA=1, nLambda=50, maxfun=300, factor=1.02, lambda_fixed=1.8, brent_lo=0.0,
                    brent_hi=10.0, brent_xatol=1e-5, log_det_L=0.0, flags=0, echo_rank=0)
    assert L.met2_t2_workspace_bytes(10, ctypes.byref(cfg)) > 0
    emu.counters(reset_only=True)
    for field, bad in (("nT2", 0), ("nT2", 129), ("nTE", 65), ("nA", 0), ("method", 6), ("method", -1)):
        c = emu.T2Cfg.from_buffer_copy(cfg)
        setattr(c, field, bad)
        assert L.met2_t2_workspace_bytes(10, ctypes.byref(c)) == -1
        assert b"met2_t2" in L.met2_last_error()
    c = emu.T2Cfg.from_buffer_copy(cfg)
    c.method, c.nLambda = 3, 2                                      # L-curve needs at least 3 lambdas
    assert L.met2_t2_workspace_bytes(10, ctypes.byref(c)) == -1
    buf = np.zeros(4096)
    p = buf.ctypes.data_as(ctypes.c_void_p)
    fn = L.met2_t2_fit
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(emu.T2Cfg)] + [ctypes.c_void_p] * 14
    args = [p, p, 0, ctypes.byref(cfg)] + [p] * 13 + [None]
    assert fn(*args) == 0                                           # V = 0: nothing to do
    args[2] = 1
    args[7] = None                                                  # X2 without kband
    assert fn(*args) == -1 and b"kband" in L.met2_last_error()
    args[7] = p
    args[4] = None                                                  # dic = NULL
    assert fn(*args) == -1 and b"NULL" in L.met2_last_error()
    fcfg = emu.FaCfg(method=1, nTE=m, nT2=n, nA=273, nKnots=3, final_solve=1, brent_lo=90.0, brent_hi=180.0,
                     brent_xatol=1e-5, brent_maxfun=500, reserved=0)
    assert L.met2_fa_workspace_bytes(10, ctypes.byref(fcfg)) == -1 and b"knots" in L.met2_last_error()
    assert emu.counters()["launches"] == 0
    d, _ = emu.epg_dictionary([120.0], [20.0, 80.0], [1000.0, 1000.0], 8, 10.0, 1000.0)
    assert d.shape == (1, 8, 2) and emu.counters()["launches"] == 1
    with pytest.raises(RuntimeError, match="bad argument"):
        emu.epg_dictionary([120.0], np.ones(200), np.ones(200), 8, 10.0, 1000.0)      # nT2 > MET2_MAX_NT2
