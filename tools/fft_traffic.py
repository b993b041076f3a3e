#!/usr/bin/env python
"""profiles/fft_traffic.json from an `ncu --set full` capture of tools/fft_probe.py: DRAM bytes (read + write) of ONE launch of each commit-transform
kernel per trace element.  bench.py multiplies it by the elements of a step for `roofline.traffic`.

    python tools/fft_traffic.py gpurun_out/ncu_fft.ncu-rep <log_rows> <n_cols> > profiles/fft_traffic.json
"""
import csv
import json
import re
import subprocess
import sys


def main():
    rep, log_rows, n_cols = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per, t_ms = {}, {}
    for r in rows[2:]:
        name = re.sub(r"\((?:bool|int|unsigned int)\)", "", r[ki])
        name = re.sub(r">\(.*$", ">", name).replace("nb::", "")
        if "fft_" not in name or name in per:
            continue                      # first launch of each kernel = one transform of the batch
        b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
        per[name] = b / (n_cols << log_rows)
        t_ms[name] = float(r[ti].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(units[ti], 1e-6)
    print(json.dumps({"dram_bytes_per_trace_element": sum(per.values()), "per_kernel": per, "kernel_ms_under_ncu": t_ms,
                      "source": f"ncu --set full, {rep} (tools/fft_probe.py {log_rows} {n_cols}): dram__bytes_read.sum + dram__bytes_write.sum of the first launch of each commit-transform kernel",
                      "algorithmic": 12}, indent=1))


if __name__ == "__main__":
    main()
