#!/usr/bin/env python
"""Per-source-line hot spots of a kernel from an .ncu-rep (read here, no GPU).

ncu's SASS page gives per-instruction executed counts and stall samples; nvdisasm -gi on the same
cubin gives the (inlined) source line of every instruction.  Joined and aggregated per source line:

  python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring (demangled, report)> [mangled-substring (cubin)] [top-N]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


HELPERS = {"lunar_core.cuh": (72, 94), "detmath.cuh": (1, 10 ** 6), "philox.cuh": (1, 10 ** 6)}


def is_helper(loc):
    r = HELPERS.get(loc[0])
    return r is not None and r[0] <= loc[1] <= r[1]


def line_table(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    # every cubin of the library (one per translation unit); `kernel` is matched against the MANGLED section name
    dis = "".join(subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout
                  for f in sorted(os.listdir(tmp)) if f.endswith(".cubin"))
    table, cur, group, infn, taken = {}, None, [], False, None
    for ln in dis.splitlines():
        if ln.startswith("//--------------------- .text."):
            infn = kernel in ln and taken in (None, ln)
            if infn:
                taken = ln
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            group.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m:
            if group:
                # innermost location that is not a one-line helper (vector algebra, detmath): the caller's line
                cur = next((g for g in group if not is_helper(g)), group[-1])
                group = []
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, lib, kernel = sys.argv[1:4]   # kernel: substring of the demangled name in the report
    mangled = sys.argv[4] if len(sys.argv) > 4 and not sys.argv[4].isdigit() else kernel   # substring of the mangled name in the cubin
    top = int(sys.argv[-1]) if sys.argv[-1].isdigit() else 40
    table = line_table(lib, mangled)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # one section per profiled launch: a "Kernel Name" row, the header row, then the instructions; take the first section
    # whose kernel name contains the pattern (all blanks removed on both sides)
    squeeze = lambda t: t.replace(" ", "").replace("(bool)", "").replace("gymcuda::", "")   # noqa: E731
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sec = next((i for i in starts if squeeze(kernel) in squeeze(rows[i][1])), starts[0] if starts else -1)
    hi = next(i for i, r in enumerate(rows) if i > sec and r and r[0] == "Address")
    end = next((i for i in starts if i > hi), len(rows))
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    base = None
    agg = {}
    tot_i = tot_s = 0
    for r in rows[hi + 1:end]:
        if len(r) < len(hdr):
            continue
        addr = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
        if base is None:
            base = addr
        off = addr - base
        loc = table.get(off, (("?", 0), ""))[0] or ("?", 0)
        inst = float(r[ix["Instructions Executed"]] or 0)
        thr = float(r[ix["Thread Instructions Executed"]] or 0)
        smp = float(r[ix["# Samples"]] or 0)
        a = agg.setdefault(loc, [0.0, 0.0, 0.0, 0])
        a[0] += inst; a[1] += thr; a[2] += smp; a[3] += 1
        tot_i += inst; tot_s += smp
    print("# %s: %.0f warp-instructions, %.0f stall samples" % (kernel, tot_i, tot_s))
    print("%-22s %6s %8s %8s %7s %6s" % ("file:line", "sass", "inst %", "samples%", "lanes", ""))
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s %6d %8.2f %8.2f %7.1f" % ("%s:%d" % loc, a[3], 100 * a[0] / tot_i, 100 * a[2] / max(tot_s, 1), a[1] / max(a[0], 1)))


if __name__ == "__main__":
    main()
