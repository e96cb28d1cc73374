"""Static SASS opcode histogram of the main kernels of libmet2.so (cuobjdump -sass): instruction count per kernel and
the most frequent opcodes, with the FP64 / FP64-tensor-core (DMMA) / shared-memory / shuffle classes summed.
usage: python tools/sass_histogram.py > profiles/r02_sass_opcode_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "multicomponent_t2_toolbox_b200", "libmet2.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, ops = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        ops[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        ops[kern][m.group(1)] += 1
want = ["echo16::t2_echo_x2_kernel<1>", "echo24::t2_echo_x2_kernel<1>", "echo16::t2_echo_tik_kernel<3, 1>",
        "echo16::t2_echo_reg_kernel<3, 2, 1>", "echo16::t2_echo_reg_kernel<5, 2, 1>", "echo24::t2_echo_reg_kernel<5, 4, 2>",
        "fa_search_thread_kernel<32>", "fa_select_kernel<2, 1>", "fa_search_kernel<2, 1>", "t2_fit_kernel<2, 1, 2>",
        "t2_fit_kernel<2, 1, 4>", "t2_fit_kernel<2, 1, 0>", "epg_dictionary_kernel", "echo_basis_kernel", "nesma"]
print("cuobjdump -sass multicomponent_t2_toolbox_b200/libmet2.so (sm_100a), static instruction counts")
for w in want:
    for k, c in ops.items():
        if w in k:
            tot = sum(c.values())
            base = collections.Counter()
            for o, n in c.items():
                base[o.split(".")[0]] += n
            cls = {"FP64 (DFMA DADD DMUL DSETP)": sum(base[o] for o in ("DFMA", "DADD", "DMUL", "DSETP")),
                   "DMMA (FP64 tensor core)": base["DMMA"], "LDS/STS": base["LDS"] + base["STS"],
                   "LDG/STG/LD/ST": base["LDG"] + base["STG"] + base["LD"] + base["ST"], "local LDL/STL": base["LDL"] + base["STL"],
                   "SHFL/VOTE/REDUX": base["SHFL"] + base["VOTE"] + base["REDUX"] + base["MATCH"],
                   "BAR/WARPSYNC": base["BAR"] + base["WARPSYNC"], "BRA/BSSY/BSYNC": base["BRA"] + base["BSSY"] + base["BSYNC"]}
            print("\n== %s\n   %d instructions; " % (k[:120], tot) + "; ".join("%s %d" % kv for kv in cls.items()))
            print("   top: " + ", ".join("%s %d" % kv for kv in base.most_common(14)))
            dm = [o for o in c if o.startswith("DMMA")]
            if dm:
                print("   tensor-core opcodes: " + ", ".join("%s x%d" % (o, c[o]) for o in dm))
