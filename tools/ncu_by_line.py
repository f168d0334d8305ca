#!/usr/bin/env python3
"""Aggregate an ncu source-page (SASS) CSV per CUDA source line.

ncu's CSV export of the source page is per SASS instruction; the line table comes from nvdisasm -gi
on the cubin inside the library (-lineinfo build), matched instruction by instruction.
usage: ncu_by_line.py <report.ncu-rep> <kernel-substring> [library.so] [top-N]"""
import csv
import collections
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cuda-vp9_b200", "libvp9b200.so")
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
    # split into kernels
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and row and row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] and len(row) == len(cur["hdr"]):
            cur["rows"].append(row)
    blk = [b for b in blocks if kern in b["name"]][0]
    hdr = blk["hdr"]
    ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
        cub = max(os.listdir(td), key=lambda f: os.path.getsize(os.path.join(td, f)))
        dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
    # find the function whose mangled name matches: use c++filt on .text.* sections
    funcs, name, lines, cur_line = {}, None, None, None
    inner, in_group = os.environ.get("INNER") == "1", False
    for l in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            name = m.group(1)
            funcs[name] = []
            cur_line = None
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m and name:
            # with -gi an instruction is preceded by one line per inlined frame, outermost call site
            # first: INNER=1 attributes to the innermost frame, default to the kernel-level line
            if inner or not in_group:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            in_group = True
            continue
        if name and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
            funcs[name].append(cur_line)
            in_group = False
    dem = {}
    for k in funcs:
        dem[k] = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    def norm(n):
        return n.replace("(bool)1", "true").replace("(bool)0", "false").replace("vp9b::", "").replace("void ", "").replace(" ", "")
    want = norm(blk["name"])
    key = None
    for k, d in dem.items():
        if norm(d) == want:
            key = k
    if key is None:
        cands = [k for k, d in dem.items() if kern in d and len(funcs[k]) == len(blk["rows"])]
        key = cands[0] if cands else [k for k, d in dem.items() if kern in d][0]
    table = funcs[key]
    if len(table) != len(blk["rows"]):
        print("# warning: %d SASS rows in the report, %d in the disassembly of %s" % (len(blk["rows"]), len(table), dem[key]))
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for i, row in enumerate(blk["rows"]):
        ln = table[i] if i < len(table) else None
        a = agg[ln]
        a[0] += float(row[ci["# Samples"]] or 0)
        a[1] += float(row[ci["Instructions Executed"]] or 0)
        a[2] += float(row[ci["Thread Instructions Executed"]] or 0)
    ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    srcs = {}
    print("# %s: %d stall samples, %d warp instructions" % (blk["name"], ts, ti))
    print("# file:line  stall%  inst%  lanes/inst  source")
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cuda-vp9_b200", "csrc")
    for ln, a in sorted(agg.items(), key=lambda kv: -(kv[1][0] / max(ts, 1) + kv[1][1] / max(ti, 1)))[:top]:
        text = ""
        if ln:
            p = os.path.join(root, ln[0])
            if p not in srcs and os.path.exists(p):
                srcs[p] = open(p).read().splitlines()
            if p in srcs and ln[1] <= len(srcs[p]):
                text = srcs[p][ln[1] - 1].strip()[:100]
        print("%-22s %5.1f %5.1f %5.1f  %s" % ("%s:%d" % ln if ln else "?", 100 * a[0] / max(ts, 1), 100 * a[1] / max(ti, 1),
                                                a[2] / max(a[1], 1), text))


if __name__ == "__main__":
    main()
