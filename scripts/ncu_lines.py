#!/usr/bin/env python
"""Stall samples per CUDA source line: joins `ncu --page source --csv` (SASS rows with
sampling counters) with `nvdisasm -g` line markers of the matching cubin.

    python scripts/ncu_lines.py report.ncu-rep cubin kernel_substring [top_n]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, cubin, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    line_of = {}
    cur, active, srcfile = None, False, None
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            active = kname in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = min(int(r[ia], 16) for r in rows[2:])
    agg = defaultdict(lambda: [0, 0, defaultdict(int)])
    total = 0
    for r in rows[2:]:
        off = int(r[ia], 16) - base
        key = line_of.get(off)
        s = int(r[isamp])
        agg[key][0] += s
        agg[key][1] += int(r[iex])
        for i in stall_cols:
            v = int(r[i])
            if v:
                agg[key][2][hdr[i]] += v
        total += s
    src = {}
    for key, (s, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if key:
            if key[0] not in src:
                try:
                    src[key[0]] = open("iivision_b200/csrc/" + key[0]).read().splitlines()
                except OSError:
                    src[key[0]] = []
            if key[1] - 1 < len(src[key[0]]):
                text = src[key[0]][key[1] - 1].strip()[:70]
        why = ",".join("%s=%d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print("%5.1f%% samples=%6d inst=%9d %s | %s | %s" % (
            100.0 * s / max(total, 1), s, ex, "%s:%d" % key if key else "?", text, why))


if __name__ == "__main__":
    main()
