#!/usr/bin/env python
"""DRAM traffic, duration and pipe utilisation of the captured kernels from `ncu --page raw --csv` exports -> profiles/<tag>_traffic.json
(usage: ncu_traffic.py out.json raw1.csv [raw2.csv ...]); bench.py's roofline.traffic reads the file."""
import csv
import json
import sys

WANT = {"gpu__time_duration.sum": "duration_ns", "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct", "smsp__inst_executed.sum": "warp_instructions",
        "launch__registers_per_thread": "registers", "lts__t_sector_hit_rate.pct": "l2_hit_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct"}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}


def main():
    out = {"kernels": {}, "source": sys.argv[2:]}
    for path in sys.argv[2:]:
        rows = list(csv.reader(open(path, errors="ignore")))
        h, u = rows[0], rows[1]
        for d in rows[2:]:
            if len(d) != len(h):
                continue
            name = d[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            e = {}
            for i, c in enumerate(h):
                if c in WANT and d[i] not in ("", "n/a"):
                    e[WANT[c]] = float(d[i].replace(",", "")) * UNIT.get(u[i], 1.0)
            out["kernels"][name] = e
    json.dump(out, open(sys.argv[1], "w"), indent=1)
    print(json.dumps(out["kernels"], indent=1))


if __name__ == "__main__":
    main()
