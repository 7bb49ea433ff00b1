"""Record the DRAM traffic of a kernel class from an `ncu --set full` capture into profiles/ncu_traffic.json
(the file bench.py's `roofline.traffic` is read from -- no number is typed into bench.py).

    python tools/ncu_traffic.py gpurun_out/x.ncu-rep C2a:tc_gemm 'k_tc_gemm.*false' profiles/r2_ncu_full_xxx.txt

Among the launches whose kernel name matches the regex, the one with the longest gpu__time_duration (the class's
heaviest member) is recorded: bytes = dram__bytes_read.sum + dram__bytes_write.sum of that one launch.  Runs here (no GPU).
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TUNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def main():
    rep, key, pattern, source = sys.argv[1:5]
    # a .csv is the raw page already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv > x.csv`: full
    # reports of a whole step exceed what gpurun copies back)
    raw = open(rep).read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    best = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if not re.search(pattern, d.get("Kernel Name", "")):
            continue
        dur = float(d["gpu__time_duration.sum"].replace(",", "")) * TUNIT.get(u["gpu__time_duration.sum"], 1.0)
        rd = float(d["dram__bytes_read.sum"].replace(",", "")) * UNIT[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"].replace(",", "")) * UNIT[u["dram__bytes_write.sum"]]
        if best is None or dur > best["duration_us"]:
            best = {"bytes": rd + wr, "read": rd, "write": wr, "duration_us": dur, "kernel": d["Kernel Name"][:120],
                    "grid": d.get("launch__grid_size", ""), "id": d.get("ID"), "source": source}
    if best is None:
        sys.exit(f"no launch matches {pattern!r}")
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[key] = best
    with open(path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(key, best)


if __name__ == "__main__":
    main()
