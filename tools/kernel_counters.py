"""Full-size per-kernel counters (ncu --metrics ..., one launch each) -> profiles/r02_kernel_counters.json, the record
bench.py cites for `roofline.traffic` / `executed_tflops` ("from": file:key).  usage: kernel_counters.py counters.csv V"""
import collections, csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, V = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn, mn, mu, mv = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Unit'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    u = r[mu]
    scale = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1e-3, 'ns': 1e-6, 'ms': 1.0, 's': 1e3}.get(u, 1.0)
    agg.setdefault(r[kn], collections.Counter())[r[mn]] += v * scale
out = {"voxels": V, "source": os.path.basename(path),
       "note": "one launch per kernel of tools/prof_one.py on the full config-2 volume under ncu --metrics (cold cache, "
               "serialised); executed FP64 flops = 2 DFMA + DADD + DMUL (thread level) + 512 per DMMA.8x8x4 warp instruction"}
def rec(names):
    c = collections.Counter()
    for k, m in agg.items():
        if any(n in k for n in names):
            c.update(m)
    if not c: return None
    fl = 2 * c['smsp__sass_thread_inst_executed_op_dfma_pred_on.sum'] + c['smsp__sass_thread_inst_executed_op_dadd_pred_on.sum'] \
        + c['smsp__sass_thread_inst_executed_op_dmul_pred_on.sum'] + 512 * c.get('sm__inst_executed_pipe_tensor_subpipe_dmma.sum', 0)
    tr = c['dram__bytes_read.sum'] + c['dram__bytes_write.sum']
    return {"kernels": [k for k in agg if any(n in k for n in names)], "kernel_ms_under_ncu": c['gpu__time_duration.sum'],
            "dram_bytes_read": c['dram__bytes_read.sum'], "dram_bytes_write": c['dram__bytes_write.sum'],
            "dram_bytes_per_voxel": tr / V, "executed_fp64_flops_per_voxel": fl / V,
            "dmma_warp_instructions": c.get('sm__inst_executed_pipe_tensor_subpipe_dmma.sum', 0),
            "warp_instructions_per_voxel": c['smsp__inst_executed.sum'] / V}
out["t2_echo_x2"] = rec(["t2_echo_x2_kernel"])
out["fa_spline"] = rec(["fa_search", "fa_select", "spline_weights", "reduce_partials"])
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_kernel_counters.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
