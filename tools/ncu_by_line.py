"""Aggregates an ncu source-page CSV (SASS level) by CUDA source line/function using nvdisasm -g line info.
usage: python tools/ncu_by_line.py <report.ncu-rep> <lib.so> [kernel-cubin-substr]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, lib = sys.argv[1], sys.argv[2]

def pick_function(rep, funcs_dis, ninst):
    """The captured kernel's SASS among the library's template instantiations: by its template arguments (raw page
    'Kernel Name' = rp_poa_kernel<G, KB, BPS>), else the function whose length is closest."""
    try:
        raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                                             text=True).stdout.splitlines()))
        name = raw[2][raw[0].index("Kernel Name")]
        m = re.search(r"(\w+)<([\d, ]+)>", name)
        if m:
            pat = m.group(1) + "I" + "".join("Li%sE" % a.strip() for a in m.group(2).split(",")) + "E"
            for k, v in funcs_dis.items():
                if pat in k:
                    return v
    except Exception:
        pass
    return min(funcs_dis.values(), key=lambda v: abs(len(v) - ninst))

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "host_mirror" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
funcs_dis, cur, cur_fn = {}, ("?", 0), None
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur_fn = m.group(1); funcs_dis[cur_fn] = []; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur_fn:
        funcs_dis[cur_fn].append((int(m.group(1), 16), cur[0], cur[1], m.group(2).strip()))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[2:]
base = int(data[0][ia], 16)
ninst = sum(1 for r in data if r[ia].startswith("0x"))
lines = pick_function(rep, funcs_dis, ninst)
byoff = {int(r[ia], 16) - base: (int(r[ii] or 0), int(r[isamp] or 0), r[1]) for r in data if r[ia].startswith("0x")}
# function ranges in poa_core.cuh
srcpath = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "racon_b200", "csrc", "poa_core.cuh")
src = open(srcpath).read().splitlines()
funcs = []
for n, l in enumerate(src, 1):
    m = re.match(r"\s*RP_DEV\s+[\w:<>\*&\s]+?\s+(\w+)\(", l)
    if m:
        funcs.append((n, m.group(1)))
def func_of(f, line):
    if f != "poa_core.cuh":
        return f
    name = "?"
    for n, nm in funcs:
        if n <= line:
            name = nm
    return name
agg = collections.defaultdict(lambda: [0, 0])
perline = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for off, f, line, text in lines:
    if off in byoff:
        i, s, _ = byoff[off]
        k = func_of(f, line)
        agg[k][0] += i; agg[k][1] += s
        perline[(f, line)][0] += i; perline[(f, line)][1] += s
        tot_i += i; tot_s += s
print("total inst %d samples %d" % (tot_i, tot_s))
for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s inst %6.2f%%  samples %6.2f%%" % (k, 100.0 * i / tot_i, 100.0 * s / max(1, tot_s)))
print("--- top lines by samples")
for (f, line), (i, s) in sorted(perline.items(), key=lambda kv: -kv[1][1])[:40]:
    txt = src[line - 1].strip() if f == "poa_core.cuh" and line <= len(src) else ""
    print("%s:%d inst %5.2f%% samples %5.2f%%  %s" % (f, line, 100.0 * i / tot_i, 100.0 * s / max(1, tot_s), txt[:90]))
