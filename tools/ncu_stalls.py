"""Stall-reason breakdown per source function from an ncu source-page CSV.
usage: python tools/ncu_stalls.py <report.ncu-rep> <lib.so> [poa_core.cuh path]"""
import collections, csv, os, re, subprocess, sys, tempfile
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

srcpath = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "racon_b200", "csrc", "poa_core.cuh")
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
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur_fn: funcs_dis[cur_fn].append((int(m.group(1), 16), cur[0], cur[1], m.group(2).strip()))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout.splitlines()))
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Address")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][ia], 16)
ninst = sum(1 for r in data if r[ia].startswith("0x"))
lines = pick_function(rep, funcs_dis, ninst)
byoff = {int(r[ia], 16) - base: r for r in data if r[ia].startswith("0x")}
src = open(srcpath).read().splitlines()
funcs = [(n, m.group(1)) for n, l in enumerate(src, 1) for m in [re.match(r"\s*RP_DEV\s+[\w:<>\*&\s]+?\s+(\w+)\(", l)] if m]
# region classification by SASS position: instructions between first/last line of a function body in dp are "dp"
def func_of(f, line):
    if f != "poa_core.cuh": return None
    name = "?"
    for n, nm in funcs:
        if n <= line: name = nm
    return name
# attribute header-file instructions (intrinsics) to the enclosing function by nearest previous poa_core.cuh instruction
agg = collections.defaultdict(lambda: collections.Counter())
last = "?"
for off, f, line, text in lines:
    k = func_of(f, line)
    if k is None: k = last
    else: last = k
    if k in ("load_row_smem", "load_row_gmem", "store_row_smem", "store_row_gmem", "perm", "swz", "swz_e", "load_band_smem", "load_band_gmem", "store_band_smem", "store_band_gmem", "perm_band", "band_elem_smem"): k = "dp"
    r = byoff.get(off)
    if not r: continue
    for c in stall_cols:
        v = int(r[c] or 0)
        if v: agg[k][hdr[c]] += v
tot = sum(sum(c.values()) for c in agg.values())
for k, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values())):
    s = sum(c.values())
    print("%-18s %5.1f%%  " % (k, 100.0 * s / tot) + "  ".join("%s=%.1f%%" % (n.replace("stall_", ""), 100.0 * v / s) for n, v in c.most_common(6)))
