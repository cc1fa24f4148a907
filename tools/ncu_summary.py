"""Summarise ncu exports into the markdown kept under profiles/ (usage: ncu_summary.py launches.csv raw.csv > out.md)."""
import csv
import sys
from collections import defaultdict


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    t = defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]
        t[name][0] += 1
        t[name][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in t.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, us) in sorted(t.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us/1000:.1f} | {100*us/tot:.1f}% |")
    print(f"\ntotal {tot/1000:.1f} us over {sum(v[0] for v in t.values())} launches")


def raw(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    h, u, data = rows[0], rows[1], rows[2:]
    idx = {c: i for i, c in enumerate(h)}
    want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
            ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %")]
    print("| kernel | " + " | ".join(w[1] for w in want) + " |\n|---|" + "---:|" * len(want))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:48]
        cells = []
        for k, _ in want:
            if k in idx:
                v = r[idx[k]]
                try:
                    v = f"{float(v.replace(',', '')):.2f}"
                except ValueError:
                    pass
                cells.append(f"{v} {u[idx[k]]}".strip())
            else:
                cells.append("-")
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] != "-":
        launches(sys.argv[1])
    if len(sys.argv) > 2:
        print()
        raw(sys.argv[2])
