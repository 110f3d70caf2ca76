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
    step = {k: tot[k] / cnt[k] for k in tot
            if any(n in k for n in ("split_kernel<0, 0, 0>", "split_prologue<0>"))}
    if len(step) == 2:
        t = [v for k, v in step.items() if "split_kernel" in k][0]
        print("headline step = " + " + ".join("%s %.1f us" % (k.split("::")[-1], v)
                                              for k, v in sorted(step.items(), key=lambda kv: kv[1]))
              + " per launch; split_kernel share %.1f%%" % (100.0 * t / sum(step.values())))


if __name__ == "__main__":
    main()
