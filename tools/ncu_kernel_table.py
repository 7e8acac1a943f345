"""Summarise an `ncu --csv --metrics ...` launch list of one forward into a per-kernel-family roofline table.
usage: python tools/ncu_kernel_table.py launches.csv [hbm_gbs] [tf32_tflops]"""
import collections, csv, json, os, re, sys
path = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm, tf32 = pk["hbm_gbs"], pk["bf16_tflops_sustained"] / 2
except Exception:
    hbm, tf32 = 6552.0, 682.2
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
by = collections.OrderedDict()
for r in rows:
    d = by.setdefault(r["ID"], {"k": r["Kernel Name"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
fam = collections.OrderedDict()
for d in by.values():
    name = re.sub(r"<.*", "", re.sub(r"^void |\(anonymous namespace\)::|<unnamed>::", "", d["k"])).split("(")[0]
    f = fam.setdefault(name, dict(n=0, us=0.0, rd=0.0, wr=0.0, tensor=0.0))
    us = d.get("gpu__time_duration.sum", 0.0) / 1e3      # ns -> us
    f["n"] += 1; f["us"] += us
    f["rd"] += d.get("dram__bytes_read.sum", 0.0); f["wr"] += d.get("dram__bytes_write.sum", 0.0)
    f["tensor"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * us
tot = sum(f["us"] for f in fam.values())
print(f"| kernel | launches | time (us, cold, serialised) | share | DRAM read+write (MB) | achieved DRAM GB/s | % of {hbm:.0f} GB/s | tensor pipe active % (time-weighted) |")
print("|---|---|---|---|---|---|---|---|")
for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
    gbs = (f["rd"] + f["wr"]) / (f["us"] * 1e-6) / 1e9 if f["us"] else 0
    print(f"| `{name}` | {f['n']} | {f['us']:.0f} | {100 * f['us'] / tot:.1f} % | {(f['rd'] + f['wr']) / 1e6:.0f} | {gbs:.0f} | {100 * gbs / hbm:.0f} % | {f['tensor'] / f['us'] if f['us'] else 0:.1f} |")
print(f"\ntotal {tot:.0f} us over {sum(f['n'] for f in fam.values())} launches (ncu serialises launches and runs them cold; shares, not absolutes, are comparable with bench.py)")
