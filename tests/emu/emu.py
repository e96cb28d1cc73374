"""TEST INFRASTRUCTURE ONLY — builds tests/emu/emu_kernels.cpp (the T2 fit kernels of csrc/ compiled by g++ against the
single-threaded SIMT emulator simt_emu.h) and calls it with numpy arrays.  Lets the `-m "not gpu"` tests run the
kernels' real source — control flow, shared-memory indexing, warp collectives, the blocked FP64-MMA factorisation —
on the CPU.  Not a fallback: nothing in the product package imports this."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "multicomponent_t2_toolbox_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libmet2_emu.so")

METHODS = {"NNLS": 0, "T2SPARC": 1, "X2": 2, "L_curve": 3, "GCV": 4, "BayesReg": 5}


class T2Cfg(ctypes.Structure):   # met2_t2_cfg of include/met2.h
    _fields_ = [("method", ctypes.c_int32), ("nTE", ctypes.c_int32), ("nT2", ctypes.c_int32), ("nA", ctypes.c_int32),
                ("nLambda", ctypes.c_int32), ("maxfun", ctypes.c_int32),
                ("factor", ctypes.c_double), ("lambda_fixed", ctypes.c_double),
                ("brent_lo", ctypes.c_double), ("brent_hi", ctypes.c_double), ("brent_xatol", ctypes.c_double),
                ("log_det_L", ctypes.c_double), ("flags", ctypes.c_int32), ("reserved", ctypes.c_int32)]


def build(force=False):
    deps = [os.path.join(HERE, f) for f in ("emu_kernels.cpp", "simt_emu.h")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "met2.h"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-I", HERE, "-I", CSRC,
           "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "emu_kernels.cpp"), "-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed on the SIMT-emulated kernels:\n" + r.stderr[-6000:])
    return LIB


def tables(Dic_3D, L):
    """Host (numpy) versions of what met2_epg_dictionary / met2_gram_tables hand to the fit kernel:
    dic [nA][nTE][nT2], dicT [nA][nT2][nTE], G [nA][nT2][nT2], kband [10][nT2]."""
    dic = np.ascontiguousarray(np.transpose(Dic_3D, (2, 0, 1)))
    dicT = np.ascontiguousarray(np.transpose(Dic_3D, (2, 1, 0)))
    G = np.ascontiguousarray(np.einsum("aec,aed->acd", dic, dic))
    n = dic.shape[2]
    K = L.T @ L
    kband = np.zeros((10, n))
    for d in range(5):
        for c in range(n):
            r = c + d - 2
            if 0 <= r < n:
                kband[d, c] = K[r, c]          # kband[d][c] = K[c + d - 2][c]
        for r in range(n):
            c = r + d - 2
            if 0 <= c < n:
                kband[5 + d, r] = L[r, c]      # kband[5 + d][r] = L[r][r + d - 2]
    return dic, dicT, G, kband


def t2_fit(sig, fa_index, Dic_3D, L, T2s, method="X2", flags=0, echo=False, lambdas=None, myelin_T2=40.0, warps=2,
           factor=1.02, lambda_fixed=1.8):
    """Run the (emulated) fit kernel.  Returns dict(fsol, est_signal, reg, maps, status, collectives)."""
    lib = ctypes.CDLL(build())
    lib.emu_counters((ctypes.c_longlong * 3)())   # reset
    sig = np.ascontiguousarray(sig, dtype=np.float64)
    fa_index = np.ascontiguousarray(fa_index, dtype=np.int32)
    V, m = sig.shape
    dic, dicT, G, kband = tables(np.asarray(Dic_3D, dtype=np.float64), np.asarray(L, dtype=np.float64))
    nA, _, n = dic.shape
    T2s = np.asarray(T2s, dtype=np.float64)
    comp = ((T2s <= myelin_T2) * 1 + ((T2s > myelin_T2) & (T2s <= 200.0)) * 2 + (T2s >= 200.0) * 4).astype(np.uint8)
    logT2 = np.ascontiguousarray(np.log(T2s))
    lam = np.ascontiguousarray(lambdas if lambdas is not None else np.zeros(1), dtype=np.float64)
    cfg = T2Cfg(method=METHODS[method], nTE=m, nT2=n, nA=nA, nLambda=len(lam), maxfun=300, factor=factor,
                lambda_fixed=lambda_fixed, brent_lo=0.0, brent_hi=10.0, brent_xatol=1e-5, log_det_L=0.0,
                flags=int(flags) | (64 if echo else 0), reserved=0)
    if method == "GCV":                      # algorithms.py:280
        cfg.brent_lo = 1e-8
    if method == "BayesReg":                 # bayesian_interpolation.py:100-101
        cfg.brent_lo, cfg.brent_hi, cfg.maxfun = 1e-8, 2.0, 200
        with np.errstate(divide="ignore"):
            cfg.log_det_L = float(np.log(np.linalg.det(np.asarray(L, dtype=np.float64))))
    out = dict(fsol=np.zeros((V, n)), est_signal=np.zeros((V, m)), reg=np.zeros(V), maps=np.zeros((V, 6)),
               status=np.zeros(V, dtype=np.uint32))
    P = ctypes.c_void_p

    def ptr(a):
        return a.ctypes.data_as(P)

    if echo:
        fn = lib.emu_t2_echo_x2
        fn.restype = ctypes.c_longlong
        fn.argtypes = [P, P, ctypes.c_longlong, ctypes.POINTER(T2Cfg)] + [P] * 11 + [ctypes.c_int]
        rc = fn(ptr(sig), ptr(fa_index), V, ctypes.byref(cfg), ptr(dic), ptr(dicT), ptr(G), ptr(kband), ptr(logT2),
                ptr(comp), ptr(out["fsol"]), ptr(out["est_signal"]), ptr(out["reg"]), ptr(out["maps"]),
                ptr(out["status"]), warps)
    else:
        fn = lib.emu_t2_fit
        fn.restype = ctypes.c_longlong
        fn.argtypes = [P, P, ctypes.c_longlong, ctypes.POINTER(T2Cfg)] + [P] * 12 + [ctypes.c_int]
        rc = fn(ptr(sig), ptr(fa_index), V, ctypes.byref(cfg), ptr(dic), ptr(dicT), ptr(G), ptr(kband), ptr(lam),
                ptr(logT2), ptr(comp), ptr(out["fsol"]), ptr(out["est_signal"]), ptr(out["reg"]), ptr(out["maps"]),
                ptr(out["status"]), warps)
    if rc < 0:
        raise RuntimeError("emulated kernel refused the configuration (%d)" % rc)
    out["collectives"] = int(rc)
    cnt = (ctypes.c_longlong * 3)()
    lib.emu_counters(cnt)
    out["counters"] = dict(fma=int(cnt[0]), smem=int(cnt[1]), syncwarp=int(cnt[2]))
    return out


class FaCfg(ctypes.Structure):   # met2_fa_cfg of include/met2.h
    _fields_ = [("method", ctypes.c_int32), ("nTE", ctypes.c_int32), ("nT2", ctypes.c_int32), ("nA", ctypes.c_int32),
                ("nKnots", ctypes.c_int32), ("final_solve", ctypes.c_int32),
                ("brent_lo", ctypes.c_double), ("brent_hi", ctypes.c_double), ("brent_xatol", ctypes.c_double),
                ("brent_maxfun", ctypes.c_int32), ("reserved", ctypes.c_int32)]


def fa_fit(sig, Dic_3D, alpha_values, Dic_3D_LR=None, alpha_values_spline=None, warps=2):
    """Run the (emulated) flip-angle stage: brute force over `alpha_values`, or the spline method when the coarse
    dictionary is given.  Returns dict(fa_index, fa_deg, km, fsol_sum, status)."""
    lib = ctypes.CDLL(build())
    lib.emu_counters((ctypes.c_longlong * 3)())
    sig = np.ascontiguousarray(sig, dtype=np.float64)
    V, m = sig.shape
    eye = np.eye(Dic_3D.shape[1])
    dic, dicT, G, _ = tables(np.asarray(Dic_3D, dtype=np.float64), eye)
    nA, _, n = dic.shape
    alphas = np.ascontiguousarray(alpha_values, dtype=np.float64)
    spline = Dic_3D_LR is not None
    P = ctypes.c_void_p

    def ptr(a):
        return a.ctypes.data_as(P) if a is not None else None

    dic_s = dicT_s = G_s = knots = None
    if spline:
        dic_s, dicT_s, G_s, _ = tables(np.asarray(Dic_3D_LR, dtype=np.float64), eye)
        knots = np.ascontiguousarray(alpha_values_spline, dtype=np.float64)
    cfg = FaCfg(method=1 if spline else 0, nTE=m, nT2=n, nA=nA, nKnots=(len(knots) if spline else 0), final_solve=1,
                brent_lo=90.0, brent_hi=180.0, brent_xatol=1e-5, brent_maxfun=500, reserved=0)
    out = dict(fa_index=np.zeros(V, dtype=np.int32), fa_deg=np.zeros(V), km=np.zeros(V), fsol_sum=np.zeros(n),
               status=np.zeros(V, dtype=np.uint32))
    fn = lib.emu_fa_fit
    fn.restype = ctypes.c_longlong
    fn.argtypes = [P, ctypes.c_longlong, ctypes.POINTER(FaCfg)] + [P] * 13 + [ctypes.c_int]
    rc = fn(ptr(sig), V, ctypes.byref(cfg), ptr(dic), ptr(dicT), ptr(G), ptr(alphas), ptr(dic_s), ptr(dicT_s), ptr(G_s),
            ptr(knots), ptr(out["fa_index"]), ptr(out["fa_deg"]), ptr(out["km"]), ptr(out["fsol_sum"]),
            ptr(out["status"]), warps)
    if rc < 0:
        raise RuntimeError("emulated FA stage refused the configuration (%d)" % rc)
    return out
