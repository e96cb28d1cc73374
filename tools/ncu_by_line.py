"""Attribute executed warp-instructions and stall samples of an ncu report to source lines.
usage: ncu_by_line.py report.ncu-rep disasm.txt(nvdisasm -g -c cubin) mangled_kernel_name"""
import collections, csv, io, re, subprocess, sys
rep, dis, kname = sys.argv[1:4]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ia = hdr.index('# Samples'); iex = hdr.index('Instructions Executed'); isrc = hdr.index('Source')
# parse disasm: sequence of (file,line) per instruction in the kernel
lines = open(dis).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.' + kname + ':')][0]
cur = ('?', 0); seq = []
inl = ''
for l in lines[start + 1:]:
    if l.startswith('//-----') or l.startswith('\t.section'): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]+\*/', l):
        seq.append(cur)
print('instrs disasm', len(seq), 'ncu', len(data))
agg_ex = collections.Counter(); agg_s = collections.Counter()
n = min(len(seq), len(data))
for (f, ln), r in zip(seq[:n], data[:n]):
    if r[iex].isdigit():
        agg_ex[(f, ln)] += int(r[iex]); agg_s[(f, ln)] += int(r[ia])
tot = sum(agg_ex.values()); tots = sum(agg_s.values())
print('total executed', tot, 'samples', tots)
srcs = {}
for (f, ln), c in agg_ex.most_common(45):
    if f not in srcs:
        try: srcs[f] = open('/root/repo/multicomponent_t2_toolbox_b200/csrc/' + f).read().split('\n')
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:70] if 0 < ln <= len(srcs[f]) else ''
    print('%5.2f%% ex %5.2f%% smp  %s:%d  %s' % (100 * c / tot, 100 * agg_s[(f, ln)] / tots, f, ln, text))
