"""Batched host API over libmet2.so: one call per stage for all voxels instead of the reference's joblib row loops.

`Met2Plan` owns the device-resident tables of one reconstruction set-up (EPG dictionaries, Gram tables, band forms of
the regularisation matrix, grids) — what motor/motor_recon_met2_real_data.py:204-277 builds on the host — and exposes

    fa_fit(signals[V, nTE])                -> FA index / angle / km / sum of spectra      (Step 2, motor...:349-373)
    t2_fit(signals[V, nTE], fa_index[V])   -> spectra, fitted signals, reg, six maps      (Steps 3+4, motor...:428-472)

PyTorch is used only for device buffers and streams.  There is no CPU path: constructing a plan without CUDA raises.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib, grids

FA_METHOD_CODE = {"brute-force": 0, "spline": 1}
REG_METHOD_CODE = {"NNLS": 0, "T2SPARC": 1, "X2": 2, "L_curve": 3, "GCV": 4, "BayesReg": 5}

ST_SKIPPED, ST_NONFINITE, ST_ITMAX, ST_SSE_ZERO, ST_NOT_PD = 1, 2, 4, 8, 16
REG_IS_LAMBDA, NO_NORMALISE, COLD_START, GCV_EVAL, FULL_START, GCV_GRID, ECHO_SPACE = 1, 2, 4, 8, 16, 32, 64   # MET2_T2_FLAG_*
# Largest residual of the reduction (max_j |d_j - U C_j| / max_j |d_j| over the flip angles, measured by
# met2_echo_basis) for which a rank of the reduced echo space is used.  24 (MET2_ECHO_RANK): the rounding of the
# dictionary's own entries (8e-16 .. 1e-15 measured).  16 (MET2_ECHO_RANK_SMALL): 4e-12 — the reference's 32-echo
# protocol measures 1.6e-12 (60 bins) / 1.9e-12 (96 bins) and reproduces the unmodified reference on 20 480 voxels with 0
# active-set disagreements and spectra within 7.8e-9, the accuracy class of the Gram-domain kernels (1.7e-9), 19 % less
# time per volume (profiles/r02_ab_echo_rank.json); the 48-echo protocol of config 4 measures 2.2e-11 and stays at rank
# 24; a dictionary that misses that bound too runs the Gram-domain kernels.
ECHO_TAIL_MAX = {24: 1e-15, 16: 4e-12}


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise _lib.Met2Error("met2: no CUDA device available — this package has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise _lib.Met2Error("met2: device must be a CUDA device, got %s" % dev)
    return dev


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f64(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)


def gaussian_smooth(data, sigma=2.0, truncate=4.0, device=None):
    """Per-echo 3-D Gaussian smoothing of data[nx, ny, nz, nt] on the GPU, bitwise equal to
    `scipy.ndimage.gaussian_filter(data[..., c], sigma, 0)` for every c (motor...:336-346).  Returns a CUDA tensor."""
    dev = _require_cuda(device)
    lib = _lib.load()
    if isinstance(data, np.ndarray):
        data = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float64))
    vol = data.to(device=dev, dtype=torch.float64).contiguous()
    if vol.dim() != 4:
        raise ValueError("data must be [nx, ny, nz, nt]")
    radius = int(truncate * float(sigma) + 0.5)
    k = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (sigma * sigma) * k ** 2)     # scipy.ndimage._filters._gaussian_kernel1d, order 0
    w = w / w.sum()
    wd = _dev_f64(w, dev)
    out = torch.empty_like(vol)
    tmp = torch.empty_like(vol)
    nx, ny, nz, nt = vol.shape
    with torch.cuda.device(dev):
        _lib.check(lib.met2_gaussian_smooth(_ptr(vol), nx, ny, nz, nt, _ptr(wd), radius, _ptr(out), _ptr(tmp), _stream()),
                   "met2_gaussian_smooth")
    return out


def nesma_filter(data, mask, half_window=6, threshold=2.5, device=None):
    """NESMA denoiser of motor/motor_recon_met2_real_data.py:305-333 on the GPU (met2_nesma_filter): every voxel with
    mask == 1 becomes the mean of the signals in its [-6, +6) window whose relative L1 distance to it is < 2.5 %.
    data[nx, ny, nz, nt], mask[nx, ny, nz]; returns a CUDA tensor like data."""
    dev = _require_cuda(device)
    lib = _lib.load()
    if isinstance(data, np.ndarray):
        data = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float64))
    if isinstance(mask, np.ndarray):
        mask = torch.as_tensor(np.ascontiguousarray(mask))
    vol = data.to(device=dev, dtype=torch.float64).contiguous()
    msk = mask.to(device=dev, dtype=torch.int32).contiguous()
    if vol.dim() != 4 or tuple(msk.shape) != tuple(vol.shape[:3]):
        raise ValueError("data must be [nx, ny, nz, nt] and mask [nx, ny, nz]")
    out = torch.empty_like(vol)
    tmp = torch.empty_like(vol)
    nx, ny, nz, nt = vol.shape
    with torch.cuda.device(dev):
        _lib.check(lib.met2_nesma_filter(_ptr(vol), _ptr(msk), nx, ny, nz, nt, int(half_window), float(threshold),
                                         _ptr(out), _ptr(tmp), _stream()), "met2_nesma_filter")
    return out


def segment_means(sig, fa_index, labels, n_seg, dictionary):
    """Per-segment mean signal and mean kernel (met2_segment_means; motor...:377-392 and
    motor_recon_met2_real_data_ROI.py:408-423).  sig[V, nTE], fa_index[V], labels[V] in [0, n_seg) (else: no segment).
    Returns (mean_signal[n_seg, nTE], mean_kernel[n_seg, nTE, nT2], counts[n_seg]) as CUDA tensors."""
    lib = _lib.load()
    dev = dictionary.dic.device
    if isinstance(sig, np.ndarray):
        sig = torch.as_tensor(np.ascontiguousarray(sig, dtype=np.float64))
    sig = sig.to(device=dev, dtype=torch.float64).contiguous()
    as_i32 = lambda a: (torch.as_tensor(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a).to(
        device=dev, dtype=torch.int32).contiguous()
    fa_index, labels = as_i32(fa_index), as_i32(labels)
    V = sig.shape[0]
    if sig.dim() != 2 or sig.shape[1] != dictionary.nTE or fa_index.shape != (V,) or labels.shape != (V,):
        raise ValueError("segment_means: sig[V, nTE], fa_index[V], labels[V] expected")
    n_seg = int(n_seg)
    mean_signal = torch.empty((n_seg, dictionary.nTE), dtype=torch.float64, device=dev)
    mean_kernel = torch.empty((n_seg, dictionary.nTE, dictionary.nT2), dtype=torch.float64, device=dev)
    counts = torch.empty(n_seg, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = torch.empty(int(lib.met2_segment_workspace_bytes(n_seg, dictionary.nA)), dtype=torch.uint8, device=dev)
        _lib.check(lib.met2_segment_means(_ptr(sig), _ptr(fa_index), _ptr(labels), V, dictionary.nTE, dictionary.nT2,
                                          dictionary.nA, n_seg, _ptr(dictionary.dic), _ptr(mean_signal),
                                          _ptr(mean_kernel), _ptr(counts), _ptr(ws), _stream()), "met2_segment_means")
    return mean_signal, mean_kernel, counts


class Dictionary:
    """Device EPG dictionary of one angle grid: dic [nA][nTE][nT2], dicT [nA][nT2][nTE], G [nA][nT2][nT2]."""

    def __init__(self, alphas, T2s, T1s, n_echoes, tau, TR, dev):
        lib = _lib.load()
        self.alphas_host = np.ascontiguousarray(alphas, dtype=np.float64)
        self.nA, self.nT2, self.nTE = len(self.alphas_host), len(T2s), int(n_echoes)
        self.alphas = _dev_f64(self.alphas_host, dev)
        t2 = _dev_f64(T2s, dev)
        t1 = _dev_f64(T1s, dev)
        self.dic = torch.empty((self.nA, self.nTE, self.nT2), dtype=torch.float64, device=dev)
        self.dicT = torch.empty((self.nA, self.nT2, self.nTE), dtype=torch.float64, device=dev)
        self.G = torch.empty((self.nA, self.nT2, self.nT2), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.met2_epg_dictionary(_ptr(self.alphas), self.nA, _ptr(t2), _ptr(t1), self.nT2, self.nTE,
                                               float(tau), float(TR), _ptr(self.dic), _ptr(self.dicT), _stream()),
                       "met2_epg_dictionary")
            _lib.check(lib.met2_gram_tables(_ptr(self.dic), self.nA, self.nTE, self.nT2, None, _ptr(self.G), None, None,
                                            _stream()), "met2_gram_tables")

    @classmethod
    def from_reference_layout(cls, Dic_3D, alphas, dev):
        """Device tables for a dictionary given on the host as Dic_3D[nTE, nT2, nA] (any forward model)."""
        lib = _lib.load()
        self = cls.__new__(cls)
        Dic_3D = np.asarray(Dic_3D, dtype=np.float64)
        if Dic_3D.ndim == 2:
            Dic_3D = Dic_3D[:, :, None]
        self.nTE, self.nT2, self.nA = Dic_3D.shape
        self.alphas_host = np.ascontiguousarray(alphas if alphas is not None else np.zeros(self.nA), dtype=np.float64)
        self.alphas = _dev_f64(self.alphas_host, dev)
        self.dic = _dev_f64(np.transpose(Dic_3D, (2, 0, 1)), dev)
        self.dicT = _dev_f64(np.transpose(Dic_3D, (2, 1, 0)), dev)
        self.G = torch.empty((self.nA, self.nT2, self.nT2), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.met2_gram_tables(_ptr(self.dic), self.nA, self.nTE, self.nT2, None, _ptr(self.G), None, None,
                                            _stream()), "met2_gram_tables")
        return self

    @classmethod
    def from_device(cls, dic, alphas=None):
        """Tables for a dictionary already on the GPU as dic[nA, nTE, nT2] (e.g. the per-segment mean kernels of
        `segment_means`)."""
        lib = _lib.load()
        self = cls.__new__(cls)
        dev = dic.device
        self.dic = dic.to(torch.float64).contiguous()
        self.nA, self.nTE, self.nT2 = self.dic.shape
        self.alphas_host = np.ascontiguousarray(alphas if alphas is not None else np.zeros(self.nA), dtype=np.float64)
        self.alphas = _dev_f64(self.alphas_host, dev)
        self.dicT = self.dic.transpose(1, 2).contiguous()
        self.G = torch.empty((self.nA, self.nT2, self.nT2), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.met2_gram_tables(_ptr(self.dic), self.nA, self.nTE, self.nT2, None, _ptr(self.G), None, None,
                                            _stream()), "met2_gram_tables")
        return self

    def echo_tables(self, R):
        """met2_echo_basis at rank R (built on first use, kept): (basis [nA, nTE, R], coef [nA, nT2, R], tail) with
        tail = the largest residual of the reduction over the flip angles, max_j |d_j - U C_j| / max_j |d_j|."""
        cache = self.__dict__.setdefault("_echo_tables", {})
        if R not in cache:
            lib = _lib.load()
            dev = self.dic.device
            basis = torch.empty((self.nA, self.nTE, R), dtype=torch.float64, device=dev)
            coef = torch.empty((self.nA, self.nT2, R), dtype=torch.float64, device=dev)
            tail = torch.empty(self.nA, dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.met2_echo_basis(_ptr(self.dic), self.nA, self.nTE, self.nT2, R, _ptr(basis),
                                               _ptr(coef), _ptr(tail), _stream()), "met2_echo_basis")
            cache[R] = (basis, coef, float(tail.max()))
        return cache[R]

    def echo_basis(self, ranks=None):
        """Reduced echo basis the echo-space kernels run on: (basis, coef, R) for the smallest rank R among `ranks`
        (default: both ranks the library is built for, 16 and 24; MET2_ECHO_RANKS=24 restricts them for A/B runs) whose
        measured residual is within ECHO_TAIL_MAX[R], or None when none is (then the Gram-domain kernels run)."""
        if ranks is None:
            lib = _lib.load()
            ranks = (int(lib.met2_echo_rank(1)), int(lib.met2_echo_rank(0)))
            if os.environ.get("MET2_ECHO_RANKS"):
                ranks = [r for r in ranks if str(r) in os.environ["MET2_ECHO_RANKS"].split(",")]
        for R in sorted(ranks):
            basis, coef, tail = self.echo_tables(R)
            if tail <= ECHO_TAIL_MAX[R]:
                return basis, coef, R
        return None

    def to_reference_layout(self):
        """Host copy in the reference's layout Dic_3D[nTE, nT2, nA] (epg/epg.py:155-162)."""
        return np.ascontiguousarray(self.dic.permute(1, 2, 0).cpu().numpy())


class Met2Plan:
    def __init__(self, n_echoes, tau, TR, reg_method="X2", reg_matrix="I", FA_method="spline", myelin_T2=40.0,
                 npc=None, n_alphas=None, T1=1000.0, device=None, lambda_reg=None, Laplac=None, Dic_3D=None,
                 Dic_3D_LR=None, alpha_values=None, alpha_values_spline=None, T2s=None, t2_flags=0, echo_space=True,
                 echo_ranks=None):
        """Tables for one reconstruction set-up.  By default everything is built like motor...:204-277 (EPG dictionary
        on the GPU); `Dic_3D` (+ `Dic_3D_LR`, `alpha_values`, `alpha_values_spline`, `T2s`, `Laplac`) lets a caller bring
        its own dictionary in the reference layout [nTE, nT2, nA] — used by the drop-in row workers and per-voxel API."""
        if reg_method not in REG_METHOD_CODE:
            raise ValueError("unknown reg_method %r" % (reg_method,))
        if FA_method not in FA_METHOD_CODE:
            raise ValueError("Error: Wrong FA_method option!")
        self.dev = _require_cuda(device)
        self.lib = _lib.load()
        self.reg_method, self.reg_matrix, self.FA_method = reg_method, reg_matrix, FA_method
        self.t2_flags = int(t2_flags)    # MET2_T2_FLAG_* applied to every t2_fit of this plan (e.g. GCV_GRID)
        # reduced-echo-space kernels wherever they apply (X2 and T2SPARC with a diagonal matrix: I, InvT2), where they
        # measured 1.4-3x faster than the Gram-domain ones (profiles/r02_ab_*); echo_space=False keeps the latter
        self.echo_space = bool(echo_space)
        # ranks of the reduced space this plan may use (None: the smallest of 16 / 24 the dictionary is exact at, see
        # ECHO_TAIL_MAX; (24,) pins the rank that is exact to the rounding of the dictionary's entries)
        self.echo_ranks = None if echo_ranks is None else tuple(int(r) for r in echo_ranks)
        self.nTE, self.tau, self.TR = int(n_echoes), float(tau), float(TR)
        if Dic_3D is not None:
            Dic_3D = np.asarray(Dic_3D, dtype=np.float64)
            if Dic_3D.ndim == 2:
                Dic_3D = Dic_3D[:, :, None]
            npc = Dic_3D.shape[1]
            if Dic_3D.shape[0] != self.nTE:
                raise ValueError("Dic_3D has %d echoes, expected %d" % (Dic_3D.shape[0], self.nTE))
        self.npc = grids.default_npc(reg_method) if npc is None else int(npc)
        self.T2s = grids.t2_grid(self.npc) if T2s is None else np.asarray(T2s, dtype=np.float64)
        self.T1s = float(T1) * np.ones_like(self.T2s)
        self.ind_m, self.ind_t, self.ind_csf = grids.compartment_masks(self.T2s, myelin_T2)
        self.alpha_values, self.alpha_spline = grids.fa_grids(FA_method, n_alphas)
        if Dic_3D is not None:
            self.alpha_values = (np.asarray(alpha_values, dtype=np.float64) if alpha_values is not None
                                 else np.zeros(Dic_3D.shape[2]))
            if len(self.alpha_values) != Dic_3D.shape[2]:
                raise ValueError("alpha_values does not match Dic_3D")
            if FA_method == "spline":
                if Dic_3D_LR is None or alpha_values_spline is None:
                    if alpha_values is not None:
                        raise ValueError("spline FA with a caller dictionary needs Dic_3D_LR and alpha_values_spline")
                    self.alpha_spline = None
                else:
                    self.alpha_spline = np.asarray(alpha_values_spline, dtype=np.float64)
        self.lambda_reg = grids.lambda_grid() if lambda_reg is None else np.asarray(lambda_reg, dtype=np.float64)
        self.Laplac = grids.reg_matrix(reg_matrix, self.T2s) if Laplac is None else np.asarray(Laplac, np.float64)
        K = self.Laplac.T @ self.Laplac
        off = np.abs(np.subtract.outer(np.arange(self.npc), np.arange(self.npc))) > 2
        if np.any(K[off] != 0.0) or np.any(self.Laplac[off] != 0.0):
            raise ValueError("regularisation matrix must be banded (|i-j| <= 2), like I, L1, L2, InvT2")
        with torch.cuda.device(self.dev):
            self.dict_lr = None
            if Dic_3D is not None:
                self.dict_hr = Dictionary.from_reference_layout(Dic_3D, self.alpha_values, self.dev)
                if FA_method == "spline" and self.alpha_spline is not None:
                    self.dict_lr = Dictionary.from_reference_layout(Dic_3D_LR, self.alpha_spline, self.dev)
                    self.knots = _dev_f64(self.alpha_spline, self.dev)
            else:
                self.dict_hr = Dictionary(self.alpha_values, self.T2s, self.T1s, self.nTE, tau, TR, self.dev)
                if FA_method == "spline":
                    self.dict_lr = Dictionary(self.alpha_spline, self.T2s, self.T1s, self.nTE, tau, TR, self.dev)
                    self.knots = _dev_f64(self.alpha_spline, self.dev)
            self.L_dev = _dev_f64(self.Laplac, self.dev)
            self.kband = torch.zeros((10, self.npc), dtype=torch.float64, device=self.dev)
            self.band_err = torch.zeros(1, dtype=torch.int32, device=self.dev)
            _lib.check(self.lib.met2_gram_tables(None, 0, self.nTE, self.npc, _ptr(self.L_dev), None, _ptr(self.kband),
                                                 _ptr(self.band_err), _stream()), "met2_gram_tables(L)")
            self.lambdas = _dev_f64(self.lambda_reg, self.dev)
            self.logT2 = _dev_f64(np.log(self.T2s), self.dev)
            comp = (self.ind_m.astype(np.uint8) | (self.ind_t.astype(np.uint8) << 1) | (self.ind_csf.astype(np.uint8) << 2))
            self.comp = torch.as_tensor(comp).to(self.dev)
        self._ws = {}

    # ------------------------------------------------------------------ configs
    def fa_cfg(self, final_solve=True):
        return _lib.FaCfg(method=FA_METHOD_CODE[self.FA_method], nTE=self.nTE, nT2=self.npc, nA=len(self.alpha_values),
                          nKnots=(len(self.alpha_spline) if self.alpha_spline is not None else 0),
                          final_solve=int(final_solve), brent_lo=90.0, brent_hi=180.0, brent_xatol=1e-5,
                          brent_maxfun=500, reserved=0)

    def t2_cfg(self, reg_method=None, flags=0, **overrides):
        method = self.reg_method if reg_method is None else reg_method
        cfg = _lib.T2Cfg(method=REG_METHOD_CODE[method], nTE=self.nTE, nT2=self.npc, nA=len(self.alpha_values),
                         nLambda=len(self.lambda_reg), maxfun=300, factor=1.02, lambda_fixed=1.8, brent_lo=0.0,
                         brent_hi=10.0, brent_xatol=1e-5, log_det_L=0.0, flags=int(flags) | self.t2_flags, echo_rank=0)
        if method == "GCV":
            cfg.brent_lo = 1e-8
        if method == "BayesReg":
            cfg.brent_lo, cfg.brent_hi, cfg.maxfun = 1e-8, 2.0, 200
            with np.errstate(divide="ignore"):
                cfg.log_det_L = float(np.log(np.linalg.det(self.Laplac)))
        if method == "X2" and np.array_equal(self.Laplac, np.eye(self.npc)):
            cfg.flags |= FULL_START
        if self.echo_space and self._diagonal_L() and not (cfg.flags & COLD_START):
            # measured on the config-2 volume (profiles/r02_ab_*): X2-I 365 -> 211 ms, X2-InvT2 262 -> 190 ms,
            # T2SPARC (96 bins) 282 -> 91 ms
            # L-curve and BayesReg (met2_t2_echo_reg_impl.cuh): see profiles/r02_ab_echo_reg.json
            if (method == "X2" and self.npc <= 64) or (method in ("T2SPARC", "L_curve", "BayesReg") and self.npc <= 128):
                cfg.flags |= ECHO_SPACE
        for k, v in overrides.items():   # e.g. factor=..., lambda_fixed=..., maxfun=...
            setattr(cfg, k, v)
        return cfg

    def _diagonal_L(self):
        L = self.Laplac
        return bool(np.all(L[~np.eye(self.npc, dtype=bool)] == 0.0) and np.all(np.diag(L) > 0.0))

    def _workspace(self, key, nbytes):
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.dev)
            self._ws[key] = ws
        return ws

    def _signals(self, sig):
        if isinstance(sig, np.ndarray):
            sig = torch.as_tensor(np.ascontiguousarray(sig, dtype=np.float64))
        sig = sig.to(device=self.dev, dtype=torch.float64).contiguous()
        if sig.dim() != 2 or sig.shape[1] != self.nTE:
            raise ValueError("signals must be [V, %d], got %s" % (self.nTE, tuple(sig.shape)))
        return sig

    # ------------------------------------------------------------------ Step 2
    def fa_fit(self, sig, final_solve=True, out=None):
        """Flip-angle estimation for all voxels.  Returns dict(fa_index int32[V], fa_deg[V], km[V], fsol_sum[nT2], status)."""
        sig = self._signals(sig)
        V = sig.shape[0]
        dev = self.dev
        if out is None:
            out = dict(fa_index=torch.empty(V, dtype=torch.int32, device=dev),
                       fa_deg=torch.empty(V, dtype=torch.float64, device=dev),
                       km=torch.empty(V, dtype=torch.float64, device=dev),
                       fsol_sum=torch.zeros(self.npc, dtype=torch.float64, device=dev),
                       status=torch.empty(V, dtype=torch.int32, device=dev))
        if V == 0:
            return out
        if self.FA_method == "spline" and self.dict_lr is None:
            raise ValueError("this plan was built without a coarse (spline) dictionary")
        cfg = self.fa_cfg(final_solve)
        with torch.cuda.device(dev):
            nbytes = self.lib.met2_fa_workspace_bytes(V, ctypes.byref(cfg))
            if nbytes < 0:
                _lib.check(-1, "met2_fa_workspace_bytes")
            ws = self._workspace("fa", nbytes)
            hr, lr = self.dict_hr, self.dict_lr
            _lib.check(self.lib.met2_fa_fit(
                _ptr(sig), V, ctypes.byref(cfg), _ptr(hr.dic), _ptr(hr.dicT), _ptr(hr.G), _ptr(hr.alphas),
                _ptr(lr.dic) if lr else None, _ptr(lr.dicT) if lr else None, _ptr(lr.G) if lr else None,
                _ptr(self.knots) if lr else None, _ptr(out["fa_index"]), _ptr(out["fa_deg"]), _ptr(out["km"]),
                _ptr(out["fsol_sum"]) if final_solve else None, _ptr(out["status"]), _ptr(ws), _stream()), "met2_fa_fit")
        return out

    # ------------------------------------------------------------------ Steps 3 + 4
    def t2_fit(self, sig, fa_index, reg_method=None, out=None, flags=0, dictionary=None, **cfg_overrides):
        """Spectrum fit + metrics.  Returns dict(fsol[V,nT2], est_signal[V,nTE], reg[V], maps[V,6], status[V]).
        `dictionary` (a `Dictionary` with this plan's nTE / nT2) replaces the plan's own kernel set for this call."""
        sig = self._signals(sig)
        V = sig.shape[0]
        dev = self.dev
        if isinstance(fa_index, np.ndarray):
            fa_index = torch.as_tensor(fa_index)
        fa_index = fa_index.to(device=dev, dtype=torch.int32).contiguous()
        if fa_index.shape != (V,):
            raise ValueError("fa_index must be [V]")
        if out is None:
            out = dict(fsol=torch.empty((V, self.npc), dtype=torch.float64, device=dev),
                       est_signal=torch.empty((V, self.nTE), dtype=torch.float64, device=dev),
                       reg=torch.empty(V, dtype=torch.float64, device=dev),
                       maps=torch.empty((V, 6), dtype=torch.float64, device=dev),
                       status=torch.empty(V, dtype=torch.int32, device=dev))
        if V == 0:
            return out
        cfg = self.t2_cfg(reg_method, flags, **cfg_overrides)
        hr = self.dict_hr if dictionary is None else dictionary
        if (hr.nTE, hr.nT2) != (self.nTE, self.npc):
            raise ValueError("dictionary shape does not match the plan")
        cfg.nA = hr.nA
        lambdas = self.lambdas
        if cfg.flags & GCV_GRID:    # GCV over the positive part of the L-curve grid (lambda_reg[1:]; lambda_reg[0] = 0)
            lambdas = self.lambdas[1:].contiguous()
            cfg.nLambda = lambdas.numel()
        red = None
        if cfg.flags & ECHO_SPACE:
            if not self._diagonal_L():
                raise ValueError("MET2_T2_FLAG_ECHO_SPACE needs a diagonal regularisation matrix (I, InvT2)")
            red = hr.echo_basis(self.echo_ranks)
            if red is None:      # dictionary not of numerical rank <= 24: Gram-domain kernels
                cfg.flags &= ~ECHO_SPACE
            else:
                cfg.echo_rank = red[2]
        with torch.cuda.device(dev):
            nbytes = self.lib.met2_t2_workspace_bytes(V, ctypes.byref(cfg))
            if nbytes < 0:
                _lib.check(-1, "met2_t2_workspace_bytes")
            ws = self._workspace("t2", nbytes)
            _lib.check(self.lib.met2_t2_fit_echo(
                _ptr(sig), _ptr(fa_index), V, ctypes.byref(cfg), _ptr(hr.dic), _ptr(hr.dicT), _ptr(hr.G),
                _ptr(self.kband), _ptr(lambdas), _ptr(self.logT2), _ptr(self.comp),
                _ptr(red[0]) if red else None, _ptr(red[1]) if red else None, _ptr(out["fsol"]),
                _ptr(out["est_signal"]), _ptr(out["reg"]), _ptr(out["maps"]), _ptr(out["status"]), _ptr(ws), _stream()),
                "met2_t2_fit_echo")
        return out

    def fit(self, sig, sig_fa=None):
        """Steps 2-4 for a batch of voxels: FA search on `sig_fa` (defaults to `sig`), spectrum fit on `sig`."""
        sig = self._signals(sig)
        fa = self.fa_fit(sig if sig_fa is None else sig_fa)
        t2 = self.t2_fit(sig, fa["fa_index"])
        return fa, t2
