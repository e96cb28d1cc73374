#!/usr/bin/env python
"""bench.py — headline benchmark of the per-voxel MET2 inverse-problem path (BASELINE.json metric).

One "step" = one pass of the hot path (flip-angle estimation + regularised NNLS fit + maps; Steps 2+3+4 of
motor/motor_recon_met2_real_data.py) over one synthetic whole-brain volume: BASELINE.json configs[1] =
96x96x60 (552 960 voxels), nTE=32, 60 T2 bins, reg_method=X2 (factor 1.02), reg_matrix=I, FA_method=spline.
With N GPUs every rank fits its own volume of that size (weak scaling, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [--gpus N] --steps K --warmup W   the reference's CPU algorithm (oracle port,
                                                                 all host cores, bounded sample per step)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fitted voxels/s (FA+reg NNLS)"
UNIT = "voxels/s"
SHAPE = (96, 96, 60)
N_ECHOES, TAU, TR = 32, 10.0, 1000.0
REG_METHOD, REG_MATRIX, FA_METHOD = "X2", "I", "spline"
WORKLOAD = "config2: synthetic brain 96x96x60 (552960 voxels), nTE=32, 60 T2, FA spline (15 knots -> 273 grid) + X2-I"

# Algorithmic (reference-formulation) flops per voxel, SURVEY.md §8(d): Lawson-Hanson QR on [D; sqrt(lambda) L]
# counted by oracle/flop_model.py (instrumented lh_nnls) on this workload; see DESIGN.md "Roofline".
F_ALG_T2_X2_I = 19.35e6     # Step 3 (X2-I: 1 plain + ~28 augmented NNLS, 892 LH outer iterations) [flop / voxel]
F_ALG_FA_SPLINE = 1.73e6    # Step 2 (16 plain NNLS)                                               [flop / voxel]
                            # (oracle/F_ALG.json, 12 voxels of this workload; range of the X2 figure 15.2-22.4 MFLOP)
FP64_PEAK_TFLOPS = 34.16    # own DFMA micro-benchmark on this pool's B200 (profiles/r01_fp64_peak_microbench.json);
                            # MEASURED_PEAKS.json has no FP64 entry
HBM_BYTES_PER_VOXEL = 32 * 8 + 4 + 60 * 8 + 32 * 8 + 8 + 6 * 8 + 4   # T2 kernel: read signal+index, write outputs


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default=None, help="override volume, e.g. 16,16,4 (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--t2-flags", type=int, default=0,
                    help="extra MET2_T2_FLAG_* bits for every T2 fit (A/B runs only, e.g. 64 = experimental echo-space "
                         "kernel); the default 0 is the measured configuration")
    ap.add_argument("--gram", action="store_true",
                    help="A/B runs only: Gram-domain T2 kernel instead of the default reduced-echo-space kernel")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, reasons, smax = [], set(), None
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                smax = float(parts[2])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(parts[1]))
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                         parts[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        if not sm:
            all_sm = []
            for ts, line in self.rows:
                parts = [p.strip() for p in line.split(",")]
                try:
                    all_sm.append(float(parts[1]))
                except (ValueError, IndexError):
                    pass
            sm = all_sm[-3:]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- CPU arms
_CPU_STATE = {}


def cpu_setup():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import multiprocessing

    import met2_oracle as O
    if "dic" not in _CPU_STATE:
        g = O._grids(REG_METHOD, REG_MATRIX, FA_METHOD, 40.0, N_ECHOES, TAU, TR)
        # the dictionary is set-up, not the per-voxel path: build it once outside the timed region
        _CPU_STATE["dic"] = O.create_Dic_3D(g["npc"], g["T2s"], g["T1s"], N_ECHOES, TAU, g["alpha_values"], TR)
        _CPU_STATE["dic_lr"] = O.create_Dic_3D(g["npc"], g["T2s"], g["T1s"], N_ECHOES, TAU, g["alpha_spline"], TR)
        _CPU_STATE["cores"] = os.cpu_count() or 1
        _CPU_STATE["pool"] = multiprocessing.Pool(_CPU_STATE["cores"])
    return O


def cpu_run(voxels, seed):
    """One bounded-sample pass of the reference algorithm on all host cores; returns (voxels fitted, seconds)."""
    O = cpu_setup()
    from multicomponent_t2_toolbox_b200.phantom import make_phantom
    nx = 96
    ny = max(1, int(voxels) // nx)
    ph = make_phantom((nx, ny, 1), n_echoes=N_ECHOES, tau=TAU, TR=TR, seed=seed, fa_mode="b1")
    t0 = time.perf_counter()
    O.recon_volume(ph["data"], ph["mask"], ph["TE_array"], TR, REG_METHOD, REG_MATRIX, FA_METHOD, 40.0,
                   num_cores=_CPU_STATE["cores"], Dic_3D=_CPU_STATE["dic"], Dic_3D_LR=_CPU_STATE["dic_lr"],
                   pool=_CPU_STATE["pool"])
    dt = time.perf_counter() - t0
    return nx * ny, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cpu_setup()
    cores = _CPU_STATE["cores"]
    nvox, dt = cpu_run(96 * 2, seed=100)                 # pilot to size the steps
    rate = nvox / dt
    total = max(1, args.steps + args.warmup)
    budget = min(20.0, 150.0 / total)                    # seconds of CPU work per step
    per_step = int(max(96, min(96 * 96, (rate * budget) // 96 * 96)))
    for w in range(args.warmup):
        cpu_run(per_step, seed=200 + w)
    t_total, v_total = 0.0, 0
    for k in range(args.steps):
        v, dt = cpu_run(per_step, seed=300 + k)
        v_total += v
        t_total += dt
    value = v_total / t_total
    sample = "%d voxels/step (96x%dx1 slab of the config-2 phantom), %d steps" % (per_step, per_step // 96, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference algorithm (oracle port of the reference's Python + SciPy "
                       "Lawson-Hanson path, multiprocessing over image rows); /root/reference itself is not on this box"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    _CPU_STATE["pool"].close()
    return 0


# ---------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from multicomponent_t2_toolbox_b200 import _lib, batched, pipeline
    from multicomponent_t2_toolbox_b200.phantom import make_phantom

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    json_fd = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at communicator creation: point fd 1 at stderr for the duration of
        # the run and keep the original stdout for the one JSON line
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    shape = SHAPE if args.shape is None else tuple(int(x) for x in args.shape.split(","))
    lib = _lib.load()

    ph = make_phantom(shape, n_echoes=N_ECHOES, tau=TAU, TR=TR, seed=2 + rank, fa_mode="b1", backend="gpu")
    sig_host = torch.as_tensor(ph["data"].reshape(-1, N_ECHOES)).pin_memory()
    V = sig_host.shape[0]
    plan = batched.Met2Plan(N_ECHOES, TAU, TR, reg_method=REG_METHOD, reg_matrix=REG_MATRIX, FA_method=FA_METHOD,
                            device=dev, t2_flags=args.t2_flags, echo_space=not args.gram)
    sig = sig_host.to(dev)
    fa_out = t2_out = None

    def step():
        nonlocal fa_out, t2_out
        fa_out = plan.fa_fit(sig, out=fa_out)
        t2_out = plan.t2_fit(sig, fa_out["fa_index"], out=t2_out)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        step()
    # ---- device-resident throughput (`value`)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    l0 = lib.met2_launch_count()
    t_wall0 = time.time()
    for k in range(args.steps):
        ev[k][0].record()
        fa_out = plan.fa_fit(sig, out=fa_out)
        ev[k][1].record()
        t2_out = plan.t2_fit(sig, fa_out["fa_index"], out=t2_out)
        ev[k][2].record()
    barrier()
    t_wall1 = time.time()
    launches = lib.met2_launch_count() - l0
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    fa_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    t2_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- end to end through the public API with host buffers (H2D + fit + D2H inside the timed region)
    pinned = {"fsol": torch.empty((V, plan.npc), dtype=torch.float64).pin_memory(),
              "est_signal": torch.empty((V, N_ECHOES), dtype=torch.float64).pin_memory(),
              "maps": torch.empty((V, 6), dtype=torch.float64).pin_memory(),
              "reg": torch.empty(V, dtype=torch.float64).pin_memory(),
              "fa_deg": torch.empty(V, dtype=torch.float64).pin_memory(),
              "fa_index": torch.empty(V, dtype=torch.int32).pin_memory(),
              "km": torch.empty(V, dtype=torch.float64).pin_memory(),
              "status": torch.empty(V, dtype=torch.int32).pin_memory(),
              "fa_status": torch.empty(V, dtype=torch.int32).pin_memory(),
              "fsol_sum": torch.empty(plan.npc, dtype=torch.float64).pin_memory()}
    sig_np = sig_host.numpy()
    pipeline.fit_voxels(plan, sig_np, pinned_out=pinned)
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        res = pipeline.fit_voxels(plan, sig_np, pinned_out=pinned)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    h2d = int(sig_host.numel() * 8)
    d2h = int(sum(t.numel() * t.element_size() for t in pinned.values()))
    mwf_mean = float(res["maps"][:, 0].mean())

    # ---- strong-scaling view of the same workload (north_star: "a 96x96x60 brain in under 1 s on 8 B200"): ONE
    #      config-2 volume (rank 0's, seed 2) cut into contiguous voxel slabs, one per rank, no collective in the fit
    one_ms = total_ms / args.steps
    if world > 1:
        ph0 = ph if rank == 0 else make_phantom(shape, n_echoes=N_ECHOES, tau=TAU, TR=TR, seed=2, fa_mode="b1",
                                                backend="gpu")
        slab = torch.as_tensor(ph0["data"].reshape(-1, N_ECHOES)[pipeline.cyclic_slab(V, rank, world)]).to(dev)
        for _ in range(3):
            plan.t2_fit(slab, plan.fa_fit(slab)["fa_index"])
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            plan.t2_fit(slab, plan.fa_fit(slab)["fa_index"])
        s1.record()
        barrier()
        one_ms = s0.elapsed_time(s1) / args.steps

    t = torch.tensor([total_ms, e2e_ms, fa_ms, t2_ms, one_ms], dtype=torch.float64, device=dev)
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms, fa_ms, t2_ms, one_ms = [float(x) for x in t.tolist()]
    launches = int(lt.item())
    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * V / (ms_per_step * 1e-3)
        t2_s = t2_ms * 1e-3
        achieved_tf = F_ALG_T2_X2_I * V / t2_s / 1e12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "t2_flags": args.t2_flags, "voxels_per_gpu": V, "l2": "inputs (142 MB) + outputs (0.44 GB) per "
                           "step exceed the 126 MB L2; no explicit flush", "stage_ms": {"fa": fa_ms, "t2": t2_ms}},
                "clocks": clocks,
                "e2e": {"value": world * V / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
                "gpu_launches": launches,
                "roofline": {"bound": "fp64", "kernel": "t2_fit_kernel", "achieved": achieved_tf,
                             "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved_tf / FP64_PEAK_TFLOPS,
                             "traffic": _ncu_traffic() if shape == SHAPE else None,
                             "algorithmic_bytes": HBM_BYTES_PER_VOXEL * V,
                             "note": "FP64 DFMA-issue bound, not HBM/tensor (SURVEY.md 8d); achieved = algorithmic "
                                     "(reference-formulation) flops/voxel x voxels / kernel time; peak = own DFMA "
                                     "microbench (MEASURED_PEAKS.json has no FP64 entry)",
                             "executed_tflops": (_ncu_executed_flops_per_voxel() or 0.0) * V / t2_s / 1e12
                             if shape == SHAPE else None,
                             "executed_note": "flops the Gram-domain, warm-started kernel really executes (ncu "
                                              "counters of the same launch): ~14x fewer than the reference "
                                              "formulation's F_alg, so FP64-pipe utilisation is ~5 % of peak; the "
                                              "kernel is latency/issue bound at 10 warps per SM (DESIGN.md 4)",
                             "hbm_achieved_gbs": HBM_BYTES_PER_VOXEL * V / t2_s / 1e9,
                             "hbm_peak_gbs": _hbm_peak()},
                "one_volume": {"ms": one_ms, "voxels": V, "gpus": world,
                               "note": "strong-scaling view: ONE config-2 volume dealt to %d rank(s) in chunks of 2048 voxels, device-"
                                       "resident, max over ranks; `value` stays the weak-scaling aggregate" % world},
                "check": {"mwf_mean": mwf_mean}}
        if world == 1 and not args.no_cpu_baseline:
            os.environ.setdefault("OMP_NUM_THREADS", "1")
            cpu_setup()
            nvox, dt = cpu_run(96 * 192, seed=2)
            line["cpu_baseline"] = {"value": nvox / dt, "unit": UNIT, "cores": _CPU_STATE["cores"], "kind": "port",
                                    "sample": "%d voxels (96x192x1 slab of the config-2 phantom), %.1f s" % (nvox, dt)}
            _CPU_STATE["pool"].close()
        if json_fd is None:
            print(json.dumps(line), flush=True)
        else:
            os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _ncu_executed_flops_per_voxel():
    """FP64 flops per voxel the t2_fit_kernel actually EXECUTES on this workload (2 x DFMA + DADD + DMUL thread
    instructions, same ncu capture as `_ncu_traffic`); None if the record is missing."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_t2_fit_v12_fullsize_counters.json")) as fh:
            return json.load(fh)["executed_fp64_flops_per_voxel"]
    except Exception:
        return None


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one t2_fit_kernel launch on this workload, from the committed
    ncu capture (profiles/r01_t2_fit_v12_fullsize_counters.json); None if the record is missing."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_t2_fit_v12_fullsize_counters.json")) as fh:
            return json.load(fh)["traffic_bytes"]
    except Exception:
        return None


def _hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)["hbm_gbs"]
    except Exception:
        return 6650.0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
