"""Per-source-line executed warp-instructions and stall samples of one kernel of an .ncu-rep (needs -lineinfo and
--import-source on at capture time).  usage: ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kre], capture_output=True, text=True).stdout
ex = collections.Counter(); smp = collections.Counter(); text = {}
fname = "?"; hdr = None; seen_kernel = None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if seen_kernel is None: seen_kernel = r[1]
        cur_kernel = r[1]; continue
    if r[0] == "Line No":
        hdr = r; ia = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); continue
    if hdr is None or cur_kernel != seen_kernel or len(r) <= ie or r[2] != "-":
        continue
    try:
        key = (fname, int(r[0]))
        ex[key] += int(r[ie]); smp[key] += int(r[ia]); text[key] = r[1].strip()[:90]
    except ValueError:
        pass
te, ts = sum(ex.values()), sum(smp.values())
print(seen_kernel, "executed warp-instructions", te, "samples", ts)
byfile = collections.Counter()
for (f, l), c in ex.items(): byfile[f] += c
print("by file:", {f: round(100 * c / te, 1) for f, c in byfile.items()})
for key, c in ex.most_common(top):
    print("%5.2f%% ex %5.2f%% smp  %s:%d  %s" % (100 * c / te, 100 * smp[key] / max(ts, 1), key[0], key[1], text[key]))
