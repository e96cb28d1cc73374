"""Round-2 parity tests on the device, all against fixtures written by the UNMODIFIED reference
(oracle/make_golden_r2.py, oracle/make_golden_config2.py, oracle/make_golden_methods.py):

  * plain Lawson-Hanson NNLS on the 60 / 96 / 100-bin grids, 20 480 voxels each (algorithms.py:55-82) and the 96-bin
    spline FA search (fa_estimation.py:35-70) — the wide grids have nearly collinear long-T2 columns;
  * BASELINE.json configs[3] sizes (nTE 48, 100 bins): brute-force FA, plain NNLS and BayesReg + InvT2
    (bayesian_interpolation.py:84-126) on 2 048 voxels, bounded by the reference's own reproducibility under a 1e-13
    relative perturbation of the signal;
  * the reduced-echo-space kernels (MET2_T2_FLAG_ECHO_SPACE) and the Gram-domain kernels for X2-I, X2-InvT2 and T2SPARC
    (96 bins), and the reduced echo basis itself.

Every test writes its measured rates to gpurun_out/parity_r2_*.json (copied to profiles/ after a run)."""
import json
import os

import numpy as np
import pytest
import torch

import met2_oracle as O
from multicomponent_t2_toolbox_b200 import batched

pytestmark = pytest.mark.gpu

REL_SPECTRUM = 1e-6
ABS_MAPS = 1e-4
ECHO = 64
OUTDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _rel_err(f, f_ref):
    scale = np.abs(f_ref).max(axis=1)
    scale[scale == 0] = 1.0
    return np.abs(f - f_ref).max(axis=1) / scale


def _record(name, rec):
    os.makedirs(OUTDIR, exist_ok=True)
    with open(os.path.join(OUTDIR, name), "w") as fh:
        json.dump(rec, fh, indent=1)


def _mwf(f, plan):
    return f[:, plan.ind_m].sum(1) / (f.sum(1) + 1e-16)


def test_plain_nnls_wide_grids_vs_reference(golden_plain_wide):
    """North-star bar "active sets bit-exact" for the plain solver on every grid the reference uses: 0 disagreements
    expected (the Gram-domain dependence test alone left 1 in 1 500 at 96 bins: dspace_candidate, met2_nnls.cuh)."""
    g = golden_plain_wide
    sig = g["sig"]
    V = sig.shape[0]
    rec = {}
    for npc in (60, 96, 100):
        plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method="NNLS", reg_matrix="I", FA_method="spline", npc=npc)
        out = plan.t2_fit(sig, g["fa_idx"])
        f, f_ref = out["fsol"].cpu().numpy(), g["f%d" % npc]
        bad = np.any((f > 0) != (f_ref > 0), axis=1)
        rel = _rel_err(f, f_ref)
        dmwf = np.abs(_mwf(f, plan) - _mwf(f_ref, plan))
        r = dict(voxels=int(V), active_set_disagreements=int(bad.sum()), disagreement_rate=float(bad.mean()),
                 spectrum_rel_err_max_agreeing=float(rel[~bad].max()), spectrum_rel_err_max_all=float(rel.max()),
                 max_abs_dMWF_all=float(dmwf.max()), status_nonzero=int((out["status"] != 0).sum()))
        if npc == 96:
            fa = plan.fa_fit(sig)
            r["fa_spline_index_mismatches"] = int((fa["fa_index"].cpu().numpy() != g["fa96_idx"]).sum())
            r["fa_km_rel_err_max"] = float(np.max(np.abs(fa["km"].cpu().numpy() - g["fa96_km"]) / g["fa96_km"]))
        rec["nnls_%d_bins" % npc] = r
    _record("parity_r2_plain_nnls_wide.json", rec)
    for key, r in rec.items():
        assert r["active_set_disagreements"] == 0, (key, r)
        assert r["spectrum_rel_err_max_all"] < REL_SPECTRUM and r["max_abs_dMWF_all"] < ABS_MAPS, (key, r)
        assert r["status_nonzero"] == 0, (key, r)
    assert rec["nnls_96_bins"]["fa_spline_index_mismatches"] == 0, rec
    assert rec["nnls_96_bins"]["fa_km_rel_err_max"] < 1e-6, rec


def test_config4_subset_vs_reference(golden_config4):
    """configs[3] sizes against 2 048 reference-fitted voxels.  FA index and plain NNLS at full tolerance.  BayesReg +
    InvT2: the reference does not reproduce ITSELF to 1e-6 there — the fixture holds its fit of the same signals
    perturbed by 1e-13 (relative), and the evidence is flat enough for that to move lambda by up to 3e-4 and the
    spectrum by up to 1e-4 in ~5 % of the voxels.  Our result is held to that reproducibility: no more voxels beyond
    1e-6 than twice the reference's own count (+ 1 % of the voxels; measured under the SIMT emulator on 256 voxels:
    ours 20, the reference against itself 13), a median within 3x of the reference's, active sets within 2 voxels of
    its own disagreement count and maps at full tolerance."""
    g = golden_config4
    sig = g["sig"]
    V = sig.shape[0]
    plan = batched.Met2Plan(48, 8.0, 1000.0, reg_method="BayesReg", reg_matrix="InvT2", FA_method="brute-force", npc=100)
    fa = plan.fa_fit(sig)
    idx_ref = g["fa_idx"].astype(np.int32)
    fa_bad = fa["fa_index"].cpu().numpy() != idx_ref
    nn = plan.t2_fit(sig, idx_ref, reg_method="NNLS")
    fn, fn_ref = nn["fsol"].cpu().numpy(), g["f_nnls"]
    nn_bad = np.any((fn > 0) != (fn_ref > 0), axis=1)
    rec = dict(voxels=int(V), fa_index_mismatches=int(fa_bad.sum()),
               nnls_active_set_disagreements=int(nn_bad.sum()), nnls_spectrum_rel_err_max=float(_rel_err(fn, fn_ref).max()))
    f_ref, f_p = g["f_bayes"], g["f_bayesp"]
    sup_self = np.any((f_p > 0) != (f_ref > 0), axis=1)
    d_self = _rel_err(f_p, f_ref)
    l_self = np.abs(g["bayesp_reg"] - g["bayes_reg"]) / g["bayes_reg"]
    dmwf_self = np.abs(_mwf(f_p, plan) - _mwf(f_ref, plan))
    # both kernel families: reduced echo space (rank 24 for this protocol; the plan's default) and Gram domain
    for key, echo in (("bayes", True), ("bayes_gram", False)):
        pk = plan if echo else batched.Met2Plan(48, 8.0, 1000.0, reg_method="BayesReg", reg_matrix="InvT2",
                                                FA_method="brute-force", npc=100, echo_space=False)
        assert bool(pk.t2_cfg().flags & ECHO) == echo
        t2 = pk.t2_fit(sig, idx_ref)
        f = t2["fsol"].cpu().numpy()
        reg = t2["reg"].cpu().numpy()
        sup_bad = np.any((f > 0) != (f_ref > 0), axis=1)
        d_ours = _rel_err(f, f_ref)
        l_ours = np.abs(reg - g["bayes_reg"]) / g["bayes_reg"]
        dmwf = np.abs(_mwf(f, plan) - _mwf(f_ref, plan))
        rec[key] = dict(active_set_disagreements=int(sup_bad.sum()), reference_self_disagreements=int(sup_self.sum()),
                        spectrum_rel_over_1e6=int((d_ours > 1e-6).sum()), reference_self_over_1e6=int((d_self > 1e-6).sum()),
                        spectrum_rel_median=float(np.median(d_ours)), reference_self_median=float(np.median(d_self)),
                        spectrum_rel_max=float(d_ours.max()), reference_self_max=float(d_self.max()),
                        lambda_rel_over_1e6=int((l_ours > 1e-6).sum()), reference_self_lambda_over_1e6=int((l_self > 1e-6).sum()),
                        lambda_rel_max=float(l_ours.max()), reference_self_lambda_max=float(l_self.max()),
                        max_abs_dMWF=float(dmwf.max()), reference_self_max_abs_dMWF=float(dmwf_self.max()),
                        status_nonzero=int((t2["status"] != 0).sum()))
        if echo:
            rec[key]["echo_rank"] = int(pk.dict_hr.echo_basis(pk.echo_ranks)[2])
    _record("parity_r2_config4_subset.json", rec)
    assert rec["fa_index_mismatches"] == 0 and rec["nnls_active_set_disagreements"] == 0, rec
    assert rec["nnls_spectrum_rel_err_max"] < REL_SPECTRUM, rec
    slack = int(0.01 * V)
    for key in ("bayes", "bayes_gram"):
        b = rec[key]
        assert b["status_nonzero"] == 0, rec
        assert b["active_set_disagreements"] <= b["reference_self_disagreements"] + 2, rec
        assert b["spectrum_rel_over_1e6"] <= 2 * b["reference_self_over_1e6"] + slack, rec
        assert b["lambda_rel_over_1e6"] <= 2 * b["reference_self_lambda_over_1e6"] + slack, rec
        assert b["spectrum_rel_median"] <= 3.0 * b["reference_self_median"] + 1e-9, rec
        assert b["spectrum_rel_max"] <= 10.0 * b["reference_self_max"], rec
        assert b["max_abs_dMWF"] < ABS_MAPS, rec


def test_invt2_methods_subset_vs_reference(golden_methods, golden_methods_invt2):
    """reg_matrix InvT2 at scale — where the echo-space formulation is worst conditioned (columns scaled by T2, up to
    2000): 2 048 voxels fitted by the unmodified reference with X2, L_curve and BayesReg (tests/golden/
    methods_invt2_subset.npz), both kernel families.  X2 and L_curve at full tolerance (the L-curve's corner must be THE
    grid value the reference picked); BayesReg bounded by the reference's own reproducibility under a 1e-13 perturbation
    of the signal, like the config-4 test."""
    g = golden_methods_invt2
    sig = g["sig"]
    V = sig.shape[0]
    idx = golden_methods["fa_spline_60"].astype(np.int32)
    rec = {}
    for method in ("X2", "L_curve", "BayesReg"):
        key = method + "_InvT2"
        f_ref, reg_ref = g["spectrum"](key), g[key + "_reg"]
        for fam, echo in (("echo", True), ("gram", False)):
            plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method=method, reg_matrix="InvT2", FA_method="spline", echo_space=echo)
            assert bool(plan.t2_cfg().flags & ECHO) == echo
            out = plan.t2_fit(sig, idx)
            f, reg = out["fsol"].cpu().numpy(), out["reg"].cpu().numpy()
            bad = np.any((f > 0) != (f_ref > 0), axis=1)
            rel = _rel_err(f, f_ref)
            dreg = np.abs(reg - reg_ref) / np.maximum(np.abs(reg_ref), 1e-300)
            dmwf = np.abs(_mwf(f, plan) - _mwf(f_ref, plan))
            r = dict(voxels=int(V), active_set_disagreements=int(bad.sum()), spectrum_rel_over_1e6=int((rel > 1e-6).sum()),
                     spectrum_rel_max_agreeing=float(rel[~bad].max()), spectrum_rel_median=float(np.median(rel)),
                     reg_rel_over_1e6=int((dreg > 1e-6).sum()), reg_rel_max=float(dreg.max()),
                     max_abs_dMWF=float(dmwf.max()), status_nonzero=int((out["status"] != 0).sum()))
            if method == "BayesReg":
                f_p, reg_p = g["spectrum"](key + "p"), g[key + "p_reg"]
                d_self = _rel_err(f_p, f_ref)
                l_self = np.abs(reg_p - reg_ref) / np.abs(reg_ref)
                r.update(reference_self_disagreements=int(np.any((f_p > 0) != (f_ref > 0), axis=1).sum()),
                         reference_self_over_1e6=int((d_self > 1e-6).sum()), reference_self_median=float(np.median(d_self)),
                         reference_self_max=float(d_self.max()), reference_self_lambda_over_1e6=int((l_self > 1e-6).sum()),
                         reference_self_lambda_max=float(l_self.max()))
            rec["%s_%s" % (key, fam)] = r
    _record("parity_r2_invt2_methods_subset.json", rec)
    slack = int(0.01 * V)
    for name, r in rec.items():
        assert r["status_nonzero"] == 0, (name, r)
        if name.startswith("BayesReg"):
            assert r["active_set_disagreements"] <= r["reference_self_disagreements"] + 2, (name, r)
            assert r["spectrum_rel_over_1e6"] <= 2 * r["reference_self_over_1e6"] + slack, (name, r)
            assert r["reg_rel_over_1e6"] <= 2 * r["reference_self_lambda_over_1e6"] + slack, (name, r)
            assert r["spectrum_rel_median"] <= 3.0 * r["reference_self_median"] + 1e-9, (name, r)
            assert r["spectrum_rel_max_agreeing"] <= 10.0 * r["reference_self_max"] + REL_SPECTRUM, (name, r)
        else:
            # X2: Brent branch points (absolute xtol 1e-5 on lambda) part two exact solvers on a few voxels in 10^4
            assert r["active_set_disagreements"] <= 2, (name, r)
            assert r["spectrum_rel_max_agreeing"] < REL_SPECTRUM, (name, r)
            assert r["spectrum_rel_over_1e6"] <= (3 if name.startswith("X2") else 0), (name, r)
            if name.startswith("L_curve"):
                assert r["reg_rel_max"] == 0.0, (name, r)      # the very grid value the reference picked
        assert r["max_abs_dMWF"] < (1e-2 if r["active_set_disagreements"] else ABS_MAPS), (name, r)


def test_echo_space_and_gram_kernels_vs_reference(golden_config2, golden_methods):
    """X2 and T2SPARC have two kernel families: the reduced-echo-space kernels (csrc/met2_t2_echo.cu; the plan's default
    for X2-I and T2SPARC, MET2_T2_FLAG_ECHO_SPACE) and the Gram-domain kernels (echo_space=False; the default for every
    other method / matrix).  Both on the device: X2-I against the 20 480 reference-fitted voxels, X2-InvT2 against the
    oracle (160 voxels) and against each other (20 480), T2SPARC (96 bins) against the 2 048 reference-fitted voxels."""
    g = golden_config2
    sig, idx = g["sig"], g["fa_idx"].astype(np.int32)
    rec = {}
    f_ref = g["f"]
    # the echo-space kernels at both ranks of the reduced space: 16 (what the plan picks for this protocol) and 24
    for name, echo, ranks in (("X2_I_echo", True, None), ("X2_I_echo_rank24", True, (24,)), ("X2_I_gram", False, None)):
        plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method="X2", reg_matrix="I", FA_method="spline", echo_space=echo,
                                echo_ranks=ranks)
        assert bool(plan.t2_cfg().flags & ECHO) == echo
        if echo:
            assert plan.dict_hr.echo_basis(ranks)[2] == (24 if ranks else 16)
        out = plan.t2_fit(sig, idx)
        f = out["fsol"].cpu().numpy()
        bad = np.any((f > 0) != (f_ref > 0), axis=1)
        rel = _rel_err(f, f_ref)
        dreg = np.abs(out["reg"].cpu().numpy() - g["reg"]) / np.abs(g["reg"])
        est = out["est_signal"].cpu().numpy()
        D = plan.dict_hr.dic.cpu().numpy()[idx[:512]]
        est_err = np.abs(est[:512] - np.einsum("vec,vc->ve", D, f[:512])).max() / np.abs(est[:512]).max()
        rec[name] = dict(voxels=int(len(f)), active_set_disagreements=int(bad.sum()),
                         spectrum_rel_err_max_agreeing=float(rel[~bad].max()),
                         k_est_rel_err_max_agreeing=float(dreg[~bad].max()),
                         max_abs_dMWF_agreeing=float(np.abs(_mwf(f, plan) - _mwf(f_ref, plan))[~bad].max()),
                         est_signal_vs_D_f_rel=float(est_err), status_nonzero=int((out["status"] != 0).sum()))
    pi = batched.Met2Plan(32, 10.0, 1000.0, reg_method="X2", reg_matrix="InvT2", FA_method="spline")
    pg = batched.Met2Plan(32, 10.0, 1000.0, reg_method="X2", reg_matrix="InvT2", FA_method="spline", echo_space=False)
    assert (pi.t2_cfg().flags & ECHO) and not (pg.t2_cfg().flags & ECHO)
    e = pi.t2_fit(sig, idx)
    d = pg.t2_fit(sig, idx)
    fe, fd = e["fsol"].cpu().numpy(), d["fsol"].cpu().numpy()
    bad_ab = np.any((fe > 0) != (fd > 0), axis=1)
    n_or = 160
    Dic = pi.dict_hr.to_reference_layout()
    f_or, _, reg_or = O.fitting_slice_T2(np.ones(n_or), sig[:n_or], idx[:n_or], n_or, Dic, pi.lambda_reg, 60, 32, "X2",
                                         pi.Laplac)
    rec["X2_InvT2"] = dict(voxels=int(len(fe)), active_set_disagreements_echo_vs_gram=int(bad_ab.sum()),
                           spectrum_rel_err_over_1e6_echo_vs_gram=int((_rel_err(fe, fd) > REL_SPECTRUM).sum()),
                           oracle_voxels=n_or,
                           echo_active_set_disagreements_vs_oracle=int(np.any((fe[:n_or] > 0) != (f_or > 0), axis=1).sum()),
                           echo_spectrum_rel_err_max_vs_oracle=float(_rel_err(fe[:n_or], f_or).max()),
                           gram_active_set_disagreements_vs_oracle=int(np.any((fd[:n_or] > 0) != (f_or > 0), axis=1).sum()),
                           gram_spectrum_rel_err_max_vs_oracle=float(_rel_err(fd[:n_or], f_or).max()),
                           status_nonzero=int((e["status"] != 0).sum() + (d["status"] != 0).sum()))
    gm = golden_methods
    idx96 = gm["fa_spline_96"].astype(np.int32)
    ft_ref = gm["spectrum"]("T2SPARC_InvT2", 96)
    for name, echo in (("T2SPARC_InvT2_96_echo", True), ("T2SPARC_InvT2_96_gram", False)):
        pt = batched.Met2Plan(32, 10.0, 1000.0, reg_method="T2SPARC", reg_matrix="InvT2", FA_method="spline", npc=96,
                              echo_space=echo)
        assert bool(pt.t2_cfg().flags & ECHO) == echo
        t = pt.t2_fit(gm["sig"], idx96)
        ft = t["fsol"].cpu().numpy()
        bad_t = np.any((ft > 0) != (ft_ref > 0), axis=1)
        rec[name] = dict(voxels=int(len(ft)), active_set_disagreements=int(bad_t.sum()),
                         spectrum_rel_err_max_agreeing=float(_rel_err(ft, ft_ref)[~bad_t].max()),
                         max_abs_dMWF=float(np.abs(_mwf(ft, pt) - _mwf(ft_ref, pt)).max()),
                         status_nonzero=int((t["status"] != 0).sum()))
    _record("parity_r2_echo_space.json", rec)
    for name in ("X2_I_echo", "X2_I_echo_rank24", "X2_I_gram"):
        r = rec[name]
        assert r["status_nonzero"] == 0 and r["active_set_disagreements"] <= 4, rec  # Brent branch points (config-2 test)
        assert r["spectrum_rel_err_max_agreeing"] < REL_SPECTRUM and r["max_abs_dMWF_agreeing"] < ABS_MAPS, rec
        # Est_Signal = D f (motor...:155): to rounding, or to the residual of the rank-16 reduction (<= 4e-12)
        assert r["est_signal_vs_D_f_rel"] < (5e-12 if name == "X2_I_echo" else 1e-12), rec
    r = rec["X2_InvT2"]
    assert r["status_nonzero"] == 0, rec
    assert r["echo_active_set_disagreements_vs_oracle"] == 0 and r["gram_active_set_disagreements_vs_oracle"] == 0, rec
    assert r["echo_spectrum_rel_err_max_vs_oracle"] < REL_SPECTRUM and r["gram_spectrum_rel_err_max_vs_oracle"] < REL_SPECTRUM, rec
    # Brent branch points (absolute xtol 1e-5 on lambda): two exact solvers part ways on ~4 voxels in 10^4, like the
    # reference against itself (config-2 test; warm/cold A/B of round 1: 15 of 552 960)
    assert r["active_set_disagreements_echo_vs_gram"] <= 4 and r["spectrum_rel_err_over_1e6_echo_vs_gram"] <= 12, rec
    for name in ("T2SPARC_InvT2_96_echo", "T2SPARC_InvT2_96_gram"):
        r = rec[name]
        assert r["status_nonzero"] == 0 and r["active_set_disagreements"] == 0, rec
        assert r["spectrum_rel_err_max_agreeing"] < REL_SPECTRUM and r["max_abs_dMWF"] < ABS_MAPS, rec


def test_lcurve_corner_echo_equals_gram_full_volume():
    """The L-curve corner (algorithms.py:88-113, 150-206) is a discrete choice that looks at neighbouring grid points:
    the reduced-echo-space kernel (default) and the Gram-domain kernel must pick the same lambda over the whole config-2
    volume.  (The first echo-space version solved the whole grid in echo space and picked a spurious corner at
    lambda <= 2.4e-8 in 531 of 552 960 voxels with InvT2, where the oracle sided with the Gram-domain kernel in all 121
    voxels examined — profiles/r02_lcurve_arbiter.json; the small end of the grid now stays in the Gram domain.)
    40 of the voxels are also checked against the oracle."""
    from multicomponent_t2_toolbox_b200.phantom import make_phantom
    ph = make_phantom((96, 96, 60), seed=2, fa_mode="b1", backend="gpu")
    sig_h = ph["data"].reshape(-1, 32)
    sig = torch.as_tensor(sig_h).cuda()
    rec = {}
    fa = None
    for rm in ("I", "InvT2"):
        reg = {}
        for name, echo in (("echo", True), ("gram", False)):
            plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method="L_curve", reg_matrix=rm, FA_method="spline", echo_space=echo)
            assert bool(plan.t2_cfg().flags & ECHO) == echo
            if fa is None:
                fa = plan.fa_fit(sig)
            out = plan.t2_fit(sig, fa["fa_index"])
            assert int((out["status"] != 0).sum()) == 0
            reg[name] = out["reg"].cpu().numpy()
            if echo:
                f_echo = out["fsol"]
            else:
                rel = float(((f_echo - out["fsol"]).abs().max(dim=1).values /
                             out["fsol"].abs().max(dim=1).values.clamp_min(1e-300))[torch.as_tensor(reg["echo"] == reg["gram"]).cuda()].max())
        differ = int((reg["echo"] != reg["gram"]).sum())
        idx = fa["fa_index"].cpu().numpy()
        Dic = plan.dict_hr.to_reference_layout()
        pick = np.random.default_rng(3).choice(len(idx), 40, replace=False)
        bad_or = 0
        for v in pick:
            _, _, reg_ref = O.t2_fit_voxel(sig_h[v], np.ascontiguousarray(Dic[:, :, idx[v]]), "L_curve", plan.Laplac, plan.lambda_reg)
            bad_or += int(reg["echo"][v] != reg_ref)
        rec["L_curve_" + rm] = dict(voxels=int(len(idx)), corner_differs_echo_vs_gram=differ,
                                    spectrum_rel_max_same_corner=rel, oracle_voxels=40, echo_corner_differs_from_oracle=bad_or)
    _record("parity_r2_lcurve_corner.json", rec)
    for key, r in rec.items():
        assert r["corner_differs_echo_vs_gram"] <= 3, rec          # <= 5e-6 of the voxels
        assert r["echo_corner_differs_from_oracle"] == 0, rec
        assert r["spectrum_rel_max_same_corner"] < REL_SPECTRUM, rec


def test_echo_basis_reproduces_the_dictionary():
    """met2_echo_basis: U orthonormal, U C = D — to rounding at rank 24 for the three dictionary shapes in use, to 4e-12
    at rank 16 for the reference's 32-echo protocol (the bound under which the plan picks rank 16, batched.ECHO_TAIL_MAX)."""
    dev = torch.device("cuda", 0)
    rec = {}
    for nte, tau, npc, alphas in ((32, 10.0, 60, np.linspace(90, 180, 273)), (32, 10.0, 96, np.linspace(90, 180, 15)),
                                  (48, 8.0, 100, np.linspace(90, 180, 91))):
        T2s = np.logspace(1, np.log10(2000.0), npc)
        d = batched.Dictionary(alphas, T2s, 1000.0 * np.ones(npc), nte, tau, 1000.0, dev)
        D = d.dic.cpu().numpy()
        tails = {}
        for R, tail_max, resid_max in ((24, 1e-15, 2e-15), (16, None, None)):
            Ut, Ct, tail = d.echo_tables(R)
            tails[R] = tail
            U, C = Ut.cpu().numpy(), Ct.cpu().numpy()
            UtU = np.einsum("ake,akf->aef", U, U)
            nz = np.abs(np.diagonal(UtU, axis1=1, axis2=2)) > 0.5       # directions below D's rounding are zero vectors
            eye = np.zeros_like(UtU)
            ii = np.arange(UtU.shape[1])
            eye[:, ii, ii] = nz
            assert np.abs(UtU - eye).max() < 1e-14
            resid = np.abs(np.einsum("ake,aje->akj", U, C) - D).max() / np.abs(D).max()
            if R == 24:
                assert tail <= tail_max and resid <= resid_max, (nte, npc, R, tail, resid)
            else:
                assert resid <= 4.0 * tail + 1e-15, (nte, npc, R, tail, resid)     # the kernel's own measure is honest
        picked = d.echo_basis()
        assert picked is not None and picked[2] == (16 if tails[16] <= batched.ECHO_TAIL_MAX[16] else 24)
        if nte == 32:
            assert picked[2] == 16, tails
        rec["nTE%d_nT2%d" % (nte, npc)] = dict(tail_rank16=tails[16], tail_rank24=tails[24], rank_picked=int(picked[2]))
    _record("parity_r2_echo_basis.json", rec)


def test_fa_search_thread_kernel_equals_warp_kernel():
    """The flip-angle search has two kernels (csrc/met2_fa.cu): thread per voxel (default) and warp per voxel, which also
    redoes the voxels the first one hands back.  fa_estimation.py:35-112 — the chosen index must be the same whichever
    path a voxel takes: the whole config-2 phantom (552 960 voxels, spline) and a 91-angle brute-force slab, each through
    the default path, the warp kernel alone and a forced hand-back of most voxels (cap of 3 columns).  Fixtures of the
    unmodified reference cover the default path elsewhere (test_config2_subset_vs_reference, ...)."""
    from multicomponent_t2_toolbox_b200.phantom import make_phantom
    ph = make_phantom((96, 96, 60), seed=2, fa_mode="b1", backend="gpu")
    sig = torch.as_tensor(ph["data"].reshape(-1, 32)).cuda()
    rec = {}
    for fam, sl in (("spline", slice(None)), ("brute-force", slice(0, 96 * 96 * 4))):
        plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method="X2", reg_matrix="I", FA_method=fam)
        x = sig[sl].contiguous()
        base = plan.fa_fit(x)
        assert int((base["status"] != 0).sum()) == 0
        for env in ({"MET2_FA_SEARCH": "warp"}, {"MET2_FA_THREAD_PCAP": "3"}):
            os.environ.update(env)
            try:
                alt = plan.fa_fit(x)
            finally:
                for k in env:
                    del os.environ[k]
            mism = int((alt["fa_index"] != base["fa_index"]).sum())
            rec["%s_%s" % (fam, "_".join("%s=%s" % kv for kv in env.items()))] = dict(voxels=int(x.shape[0]), index_mismatches=mism)
            assert mism == 0, (fam, env, mism)
            assert torch.equal(alt["status"], base["status"])
            assert float(((alt["km"] - base["km"]).abs() / base["km"].clamp_min(1e-300)).max()) < 1e-12
    _record("parity_r2_fa_search_paths.json", rec)
