#!/usr/bin/env python
"""Per-kernel launch count, total time and share from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv --log-file launches.csv ...`).

    python scripts/launch_shares.py gpurun_out/launches_final.csv
"""
import csv
import sys
from collections import defaultdict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
        name = r[ik].split("(")[0]
        tot[name] += float(r[iv].replace(",", "")) * scale
        cnt[name] += 1
    total = sum(tot.values())
    for name in sorted(tot, key=tot.get, reverse=True)[:16]:
        print("%-64s %5d %12.1f us %5.1f%%" % (name[-64:], cnt[name], tot[name],
                                               100.0 * tot[name] / total))
    pair = {k: v for k, v in tot.items() if "tree_kernel<0, 0, 0>" in k or "pixel_prologue<0>" in k}
    if len(pair) == 2:
        t = [v / cnt[k] for k, v in pair.items() if "tree_kernel" in k][0]
        p = [v / cnt[k] for k, v in pair.items() if "pixel_prologue" in k][0]
        print("headline step = pixel_prologue<0> + tree_kernel<0,0,0>: %.1f us + %.1f us per launch; "
              "tree_kernel share %.1f%%" % (p, t, 100.0 * t / (t + p)))


if __name__ == "__main__":
    main()
