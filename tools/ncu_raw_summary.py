#!/usr/bin/env python
"""Markdown table from `ncu -i X.ncu-rep --page raw --csv` (argv[1]): one row per captured launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "DRAM read GB"),
        ("dram__bytes_write.sum", "DRAM write GB"), ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "registers"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %")]
idx = [(H.index(n), lab) for n, lab in want if n in H]
units = rows[1]
print("| " + " | ".join(l for _, l in idx) + " |")
print("|" + "---|" * len(idx))
def conv(i, v):
    u = units[i]
    try: x = float(v.replace(",", ""))
    except ValueError: return v.split("(")[0].replace("void pc::", "").strip()[:44]
    if H[i] == "gpu__time_duration.sum":
        x = x / 1e6 if u.startswith("ns") else x / 1e3 if u.startswith("us") else x
        return "%.3f" % x
    if H[i].startswith("dram__bytes"):
        m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return "%.3f" % (x * m / 1e9)
    return "%.1f" % x if "." in v else v
PEAK = 6533.8                                                  # measured HBM copy bandwidth, GB/s (MEASURED_PEAKS.json)
for r in rows[2:]:
    cells = [conv(i, r[i]) for i, _ in idx]
    labs = [l for _, l in idx]
    if "DRAM % of peak" in labs and not cells[labs.index("DRAM % of peak")].strip():
        try:
            gb = float(cells[labs.index("DRAM read GB")]) + float(cells[labs.index("DRAM write GB")])
            cells[labs.index("DRAM % of peak")] = "%.1f" % (100.0 * gb / (float(cells[labs.index("ms")]) * 1e-3) / PEAK)
        except ValueError:
            pass
    print("| " + " | ".join(cells) + " |")
