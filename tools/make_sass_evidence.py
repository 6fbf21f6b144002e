#!/usr/bin/env python
"""Regenerates the 'after' SASS listings under profiles/r02/ from the library as built (no GPU needed):

  profiles/r02/sass_dp_row_loop_full_matrix_after.txt   single-chunk copy of the DP row loop (dp_rows<false>)
  profiles/r02/sass_static_after_gpu_budget_walk.txt     the 'AFTER' part: the walk loop of traceback<false>

The loops are found through the source lines of their loop statements (tools/sass_by_line.py --where), so the listings
follow the code when it moves.  The hand-written headers of the two files are kept."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "sass_by_line.py")
SRC = os.path.join(ROOT, "racon_b200", "csrc", "poa_core.cuh")


def tool(*args):
    return subprocess.run([sys.executable, TOOL] + list(args), stdout=subprocess.PIPE, text=True, check=True).stdout


def line_of(pattern):
    for n, l in enumerate(open(SRC), 1):
        if pattern in l:
            return n
    raise SystemExit("pattern not found in poa_core.cuh: " + pattern)


def clusters(line):
    out = []
    for l in tool("--where", "poa_core.cuh:%d" % line).splitlines()[1:]:
        a, _, b = l.partition("-")
        out.append((int(a, 16), int(b, 16)))
    return out


def listing(lo, hi):
    return "\n".join(tool("--range", "%x-%x" % (lo, hi)).splitlines()[1:])


def main():
    # (a) the row-block loop `for (i0 ...)`: two copies (dp_rows<true> first, dp_rows<false> second).  A copy starts at its
    # i_last computation (VIADDMNMX.U32 ..., 0x1f, ...) and ends with the statement's back edge = the end of the last cluster
    tops = tool("--grep", r"VIADDMNMX\.U32.*0x1f,").splitlines()[1:]
    dp_lo = int(re.search(r"/\*([0-9a-f]+)\*/", tops[-1]).group(1), 16)
    dp_hi = clusters(line_of("for (uint32_t i0 = 1; i0 <= nrows; i0 += G) {"))[-1][1]
    p = os.path.join(ROOT, "profiles", "r02", "sass_dp_row_loop_full_matrix_after.txt")
    head = open(p).read().split("\n\n")[0]
    body = listing(dp_lo, dp_hi)
    head = re.sub(r"\(\d+ instructions in the listing", "(%d instructions in the listing" % len(body.splitlines()), head)
    open(p, "w").write(head + "\n\n" + body + "\n")
    # (b) the walk loop `while (i != 0)`: the last cluster is traceback<false> (the banded instantiation comes first)
    cl = clusters(line_of("while (i != 0) {"))
    lo, hi = cl[-1]
    p = os.path.join(ROOT, "profiles", "r02", "sass_static_after_gpu_budget_walk.txt")
    txt = open(p).read()
    i = txt.index("=== AFTER")
    j = txt.index("\n/*", i)
    open(p, "w").write(txt[:j] + "\n" + listing(lo, hi + 0x30) + "\n")
    print("dp loop 0x%x-0x%x, walk loop 0x%x-0x%x" % (dp_lo, dp_hi, lo, hi))


if __name__ == "__main__":
    main()
