"""Copy the raw gpurun_out/ captures of one round-2 GPU call into the committed records under profiles/:
    python tools/records_r2.py r12 [HEAD]
bench line, launch-list summary, other-config timings, ncu --set full summary + per-function tables, GPU pytest tail,
parity records written by the GPU tests."""
import collections, csv, glob, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
head = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
G = lambda n: os.path.join(ROOT, "gpurun_out", n)
P = lambda n: os.path.join(ROOT, "profiles", n)
if os.path.exists(G(tag + "_bench.json")):
    line = open(G(tag + "_bench.json")).read().strip().splitlines()[-1]
    json.dump(json.loads(line), open(P("r02_bench_%s.json" % tag), "w"), indent=1)
if os.path.exists(G(tag + "_launches.csv")):
    rows = list(csv.reader(open(G(tag + "_launches.csv"))))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]; kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv: continue
        try: v = float(r[mv].replace(',', ''))
        except ValueError: continue
        ms = v / 1e6 if r[mu] == 'ns' else (v / 1e3 if r[mu] == 'us' else v)
        a = agg.setdefault(r[kn], [0, 0.0]); a[0] += 1; a[1] += ms
    tot = sum(a[1] for a in agg.values())
    out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 3 --no-cpu-baseline (HEAD %s, one B200)" % head,
           "kernel | launches | total ms | ms/launch | share of all captured kernel time"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-70s %4d %10.3f %10.4f %6.3f" % (k[:70], n, ms, ms / n, ms / tot))
    out.append("total %.3f ms" % tot)
    open(P("r02_launches_%s_summary.txt" % tag), "w").write("\n".join(out) + "\n")
if os.path.exists(G(tag + "_configs.json")):
    shutil.copy(G(tag + "_configs.json"), P("r02_other_configs_%s.json" % tag))
if os.path.exists(G(tag + "_pytest.log")):
    open(P("r02_gpu_pytest_%s.txt" % tag), "w").write("HEAD %s, python -m pytest tests -m gpu -q on one B200\n" % head + "".join(open(G(tag + "_pytest.log")).readlines()[-6:]))
rep = G(tag + "_prof.ncu-rep")
if os.path.exists(rep):
    keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed.sum', 'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores', 'sass__inst_executed_global_loads',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
            'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
            'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'smsp__thread_inst_executed_per_inst_executed.ratio',
            'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active']
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    txt = ["ncu --set full --clock-control none --import-source on, tools/prof_one.py (96x96x6 slab = 55 296 voxels of the config-2 phantom, FA spline + X2-I), HEAD %s, one B200" % head]
    names = []
    for vals in rows[2:]:
        kname = vals[hdr.index('Kernel Name')]
        names.append(kname)
        txt.append("== %s" % kname)
        for h, u, v in zip(hdr, units, vals):
            if h in keys or ('issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(v or 0) > 0.05):
                txt.append('  %-88s %-14s %s' % (h, u, v))
    txt.append("")
    txt.append("---- executed warp-instructions / stall samples per source function (tools/ncu_funcs.py) and per line (tools/ncu_lines.py)")
    for kre in ("t2_echo", "fa_search_thread", "fa_select"):
        if any(kre in n for n in names):
            txt.append(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_funcs.py"), rep, kre], capture_output=True, text=True).stdout)
            txt.append(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, kre, "25"], capture_output=True, text=True).stdout)
    open(P("r02_config2_kernels_ncu_summary_%s.txt" % tag), "w").write("\n".join(txt) + "\n")
for f in glob.glob(G("parity_*.json")):
    shutil.copy(f, P("r02_" + os.path.basename(f)))
print("records of", tag, "written")
