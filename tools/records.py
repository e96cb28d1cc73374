"""Turn the raw gpurun_out/ captures of one kernel version into the committed records under profiles/:
    python tools/records.py v12
launch list summary, full-size counters JSON, ncu --set full summary, bench line."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = lambda n: os.path.join(ROOT, "gpurun_out", n)
P = lambda n: os.path.join(ROOT, "profiles", n)


def table(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    return rows[hi], rows[hi + 1:]


h, rows = table(G("launches_%s.csv" % tag))
kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(',', ''))
    except ValueError:
        continue
    ms = v / 1e6 if r[mu] == 'ns' else (v / 1e3 if r[mu] == 'us' else v)
    a = agg.setdefault(r[kn], [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
out = ["ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 --no-cpu-baseline  "
       "(round 1, kernel %s)" % tag]
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-72s n=%3d total=%10.3f ms share=%.4f" % (k[:72], n, ms, ms / tot))
open(P("r01_launches_%s_summary.txt" % tag), "w").write("\n".join(out) + "\n")
print("\n".join(out[:6]))

h, rows = table(G("t2_traffic_%s.csv" % tag))
mn, mv, mu = h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
m = {r[mn]: (float(r[mv].replace(',', '')), r[mu]) for r in rows if len(r) > mv}
by = lambda name: m[name][0] * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(m[name][1], 1)
V = 552960
t = m['gpu__time_duration.sum']
rec = dict(kernel="t2_fit_kernel<2,1,X2>", launch="tools/prof_one.py config 2 (552960 voxels), kernel %s" % tag,
           dram_bytes_read=by('dram__bytes_read.sum'), dram_bytes_write=by('dram__bytes_write.sum'),
           traffic_bytes=by('dram__bytes_read.sum') + by('dram__bytes_write.sum'), algorithmic_bytes=1056 * V,
           executed_fp64_flops_vector=2 * m['smsp__sass_thread_inst_executed_op_dfma_pred_on.sum'][0]
           + m['smsp__sass_thread_inst_executed_op_dadd_pred_on.sum'][0]
           + m['smsp__sass_thread_inst_executed_op_dmul_pred_on.sum'][0],
           warp_instructions=m['smsp__inst_executed.sum'][0],
           kernel_ms_under_ncu=t[0] / (1e6 if t[1] == 'ns' else 1e3 if t[1] == 'us' else 1))
rec['executed_fp64_flops_per_voxel'] = rec['executed_fp64_flops_vector'] / V
rec['warp_instructions_per_voxel'] = rec['warp_instructions'] / V
json.dump(rec, open(P("r01_t2_fit_%s_fullsize_counters.json" % tag), "w"), indent=1)
print(rec)

title = "t2_fit %s (X2 kernel), %s" % (tag, sys.argv[2] if len(sys.argv) > 2 else "55296 voxels")
s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), G("prof_t2_%s.ncu-rep" % tag), title],
                   capture_output=True, text=True).stdout
open(P("r01_t2_fit_%s_ncu_summary.txt" % tag), "w").write(s)
line = [l for l in open(G("bench_%s.json" % tag)) if l.startswith("{")][-1]
open(P("r01_bench_%s.json" % tag), "w").write(line)
d = json.loads(line)
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"], d["config"]["stage_ms"], d["roofline"]["frac"],
      d.get("cpu_baseline"))
