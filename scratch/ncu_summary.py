"""Summarise an `ncu --csv` launch list (one row per kernel launch) per kernel name -> markdown table.

    python scratch/ncu_summary.py gpurun_out/launches.csv [title] > profiles/rNN_....md

Columns used when present: gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed, gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed."""
import csv
import re
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    lines = [l for l in open(path, newline="") if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    body = [r for r in rows[hdr + 1:] if len(r) == len(h)]
    if "Metric Name" in h:                      # long format: one row per (launch, metric)
        kid, kn, mn, mv, mu = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
        launches = {}
        for r in body:
            d = launches.setdefault(r[kid], {"name": r[kn]})
            v = num(r[mv])
            u = r[mu]
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            d[r[mn]] = v * scale
        recs = list(launches.values())
    else:
        units = rows[hdr + 1]
        recs = []
        for r in body[1:]:
            d = {"name": r[h.index("Kernel Name")]}
            for i, n in enumerate(h):
                if "__" in n:
                    u = units[i]
                    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
                    d[n] = num(r[i]) * scale
            recs.append(d)
    agg = {}
    for d in recs:
        nm = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("c2dsr::", "")
        nm = re.sub(r"at::native::\(anonymous namespace\)::|at::native::", "at::", nm)[:86]
        a = agg.setdefault(nm, dict(n=0, us=0.0, rd=0.0, wr=0.0, tp=0.0, dp=0.0))
        t = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["us"] += t
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tp"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * t
        a["dp"] += d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0) * t
    tot = sum(a["us"] for a in agg.values())
    print(f"### {title}: {sum(a['n'] for a in agg.values())} launches, {tot:.0f} us serialised (ncu replay: cold cache, no overlap)\n")
    print("| kernel | launches | us | share | DRAM MB (rd+wr) | tensor pipe % | DRAM % |")
    print("|---|---|---|---|---|---|---|")
    for nm, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        if a["us"] < 0.004 * tot:
            continue
        t = max(a["us"], 1e-9)
        print(f"| `{nm}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {(a['rd'] + a['wr']) / 1e6:.0f} | {a['tp'] / t:.0f} | {a['dp'] / t:.0f} |")


if __name__ == "__main__":
    main()
