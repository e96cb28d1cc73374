"""Executed warp-instructions and stall samples of one kernel of an .ncu-rep per source FUNCTION (line ranges taken from
the current csrc files: a function starts at a line matching `__device__ ... name(` or `auto name = [&]`).
usage: ncu_funcs.py report.ncu-rep kernel_regex"""
import collections, csv, io, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kre = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kre], capture_output=True, text=True).stdout
ex = collections.Counter(); smp = collections.Counter()
fname = "?"; hdr = None; seen = None; cur = None
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if seen is None: seen = r[1]
        cur = r[1]; continue
    if r[0] == "Line No": hdr = r; ia = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); continue
    if hdr is None or cur != seen or len(r) <= ie or r[2] != "-": continue
    try: ex[(fname, int(r[0]))] += int(r[ie]); smp[(fname, int(r[0]))] += int(r[ia])
    except ValueError: pass
starts = {}
pat = re.compile(r'^\s*(?:template.*>\s*)?(?:__device__|__global__|static|inline|__host__).*?\b([A-Za-z_0-9]+)\s*\(|^\s*auto\s+([A-Za-z_0-9]+)\s*=\s*\[')
for f in set(k[0] for k in ex):
    path = os.path.join(ROOT, "multicomponent_t2_toolbox_b200", "csrc", f)
    if not os.path.exists(path): continue
    L = []
    for i, line in enumerate(open(path), 1):
        m = pat.match(line)
        if m and not line.strip().startswith("//"): L.append((i, m.group(1) or m.group(2)))
    starts[f] = L
def func(f, ln):
    name = f
    for s, n in starts.get(f, []):
        if s <= ln: name = n
        else: break
    return name
fe = collections.Counter(); fs = collections.Counter()
for k, c in ex.items(): fe[func(*k)] += c; fs[func(*k)] += smp[k]
te, ts = sum(fe.values()), sum(fs.values())
print(seen, "executed warp-instructions", te, "samples", ts)
for n, c in fe.most_common(30): print("%6.2f%% ex %6.2f%% smp  %s" % (100 * c / te, 100 * fs[n] / max(ts, 1), n))
