"""`ncu -i <rep> --page raw --csv` -> one markdown row per captured launch with the metrics the docs quote.

    python scratch/ncu_full_table.py gpurun_out/r02_full_k4a.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import re
import subprocess
import sys

COLS = [("time us", "gpu__time_duration.sum", 1e-3), ("DRAM read MB", "dram__bytes_read.sum", None),
        ("DRAM write MB", "dram__bytes_write.sum", None),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
        ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("L2 hit %", "lts__t_sector_hit_rate.pct", 1), ("regs", "launch__registers_per_thread", 1),
        ("grid", "launch__grid_size", 1)]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name.replace("c2dsr::", "").replace("void ", "")[:90]


def main():
    print("| kernel | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(h, r))
            u = dict(zip(h, units))
            cells = []
            for label, key, scale in COLS:
                v = d.get(key, "")
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    cells.append(v)
                    continue
                unit = u.get(key, "")
                if scale is None:                          # bytes -> MB / KB by the unit ncu printed
                    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
                    x = x * mult / (1e3 if "KB" in label else 1e6)
                elif key.startswith("gpu__time"):
                    x = x * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1e-3)
                cells.append(f"{x:.1f}" if x < 1e5 and x != int(x) else f"{int(x)}")
            print(f"| `{short(d['Kernel Name'])}` | " + " | ".join(cells) + " |")


main()
