"""profiles/dram_traffic.json from an ncu --set full raw CSV of one sweep's launches (scripts/gpu_l.sh):

    python scripts/ncu_traffic.py <raw.csv> <key> <launches per sweep> [--out profiles/dram_traffic.json]

Sums dram__bytes_read.sum + dram__bytes_write.sum over the LAST `launches per sweep` captured kernels (one full sweep) and
records the per-kernel breakdown next to it."""
import csv, json, os, sys

raw, key, per_sweep = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "dram_traffic.json")
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: k for k, n in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}


def val(r, name):
    return float(r[col[name]].replace(",", "")) * scale.get(units[col[name]], 1)


kern = []
for r in data[-per_sweep:]:
    kern.append({"kernel": r[col["Kernel Name"]].split("(")[0][-60:], "grid": r[col["Grid Size"]],
                 "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                 "seconds_under_ncu": val(r, "gpu__time_duration.sum"),
                 "l2_hit_pct": float(r[col["lts__t_sector_hit_rate.pct"]]), "warps_active_pct": float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]])})
total = sum(k["dram_read"] + k["dram_write"] for k in kern)
d = json.load(open(out)) if os.path.exists(out) else {}
d[key] = {"bytes_per_step": total, "source": f"ncu --set full capture {os.path.basename(raw)} (profiles/)", "kernels": kern}
json.dump(d, open(out, "w"), indent=1)
print(f"{key}: {total / 1e9:.2f} GB per sweep over {len(kern)} launches")
for k in kern:
    print(f"  {k['kernel'][:50]:50s} read {k['dram_read'] / 1e9:6.2f} GB write {k['dram_write'] / 1e9:5.2f} GB  L2 hit {k['l2_hit_pct']:.0f}%  {k['seconds_under_ncu'] * 1e3:.3f} ms")
