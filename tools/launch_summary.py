#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`).

    python tools/launch_summary.py gpurun_out/launches.csv "the command that was profiled" > profiles/launches_rNN_summary.txt
"""
import collections
import csv
import re
import sys


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(r[ui], 1e-6)
        name = re.sub(r"\((?:bool|int|unsigned int)\)", "", r[ki])
        name = re.sub(r"^void ", "", name).replace("nb::", "")
        name = re.sub(r"\(.*$", "", name) if "<" not in name else re.sub(r">\(.*$", ">", name)
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {cmd}")
    print(f"# {sum(cnt.values())} launches, {total:.1f} ms of kernel time (per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)")
    print("       ms  share launches  kernel")
    for name, ms in tot.most_common():
        print(f"{ms:9.3f} {100 * ms / total:5.1f}% {cnt[name]:8d}  {name[:120]}")


if __name__ == "__main__":
    main()
