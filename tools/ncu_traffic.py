#!/usr/bin/env python
"""DRAM traffic per launch of the dominant kernels from committed-session `ncu --set full` reports -> profiles/ncu_traffic.json, the
file bench.py reads `roofline.traffic` from (so the number follows the capture instead of being a constant in the source).
usage: python tools/ncu_traffic.py render gpurun_out/r02_prof_field.ncu-rep [train gpurun_out/...]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    args = sys.argv[1:]
    for workload, path in zip(args[0::2], args[1::2]):
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        seen = set()
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").split("::")[-1]
            base = re.sub(r"<.*", "", name)
            if base in seen:          # the first launch of a kernel family in the report is the dominant variant
                continue
            seen.add(base)
            b = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(k)
                b += float(r[i].replace(",", "")) * UNIT[units[i]]
            out[f"{workload}:{base}"] = {"dram_bytes": b, "kernel": name, "ms_under_ncu": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) *
                                         {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]],
                                         "source": f"ncu --set full, {os.path.basename(path)} (summary: profiles/r02_ncu_full.md)"}
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
