"""profiles/ summaries from the GPU captures (read here, no GPU):
  python scripts/make_profile_summary.py <tag> <estep_full.ncu-rep> <launches.csv> [outdir = profiles/]
writes <outdir>/<tag>_ncu_estep_raw.csv (ncu --page raw of the whole-E-step capture), profiles/<tag>_ncu_summary.json
(per launch + whole E-step: time, DRAM bytes, pipe / issue / stall counters) and profiles/<tag>_launch_summary.json."""
import collections, csv, json, os, re, subprocess, sys

tag, rep, launches = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_ncu_estep_raw.csv"), "w").write(raw)
dig = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_digest.py"), rep], capture_output=True, text=True).stdout
mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
rows, tot_us, tot_b = [], 0.0, 0.0
for l in dig.splitlines():
    d = json.loads(l)
    d["dram_bytes"] = d.pop("dram_rd") * mul[d.pop("dram_rd_unit")] + d.pop("dram_wr") * mul[d.pop("dram_wr_unit")]
    d["name"] = re.sub(r"^void ", "", d["name"]).split("(")[0].replace("titan::", "")
    rows.append(d)
    tot_us += d["us"]
    tot_b += d["dram_bytes"]
per = collections.OrderedDict()
for d in rows:
    a = per.setdefault(d["name"], {"launches": 0, "us": 0.0, "dram_bytes": 0.0, "inst": 0.0})
    a["launches"] += 1; a["us"] += d["us"]; a["dram_bytes"] += d["dram_bytes"]; a["inst"] += d["inst"]
for a in per.values():
    a["share_of_estep"] = a["us"] / tot_us
json.dump({"what": "one whole E-step (1 M SNPs, K = 75, library defaults) under ncu --set full --clock-control none: serialised, cold-cache "
                   "launch times -- compare shares, not absolutes",
           "estep": {"us_sum_under_ncu": tot_us, "dram_bytes_per_estep": tot_b, "launches": len(rows)},
           "per_kernel": per, "launches": rows}, open(os.path.join(P, f"{tag}_ncu_summary.json"), "w"), indent=1)

rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]
i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.OrderedDict(), 0.0
for r in rows[1:]:
    n = re.sub(r"^void ", "", r[i_name]).split("(")[0]
    v = float(r[i_val].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[r[i_unit]]
    a = agg.setdefault(n, {"launches": 0, "us": 0.0})
    a["launches"] += 1; a["us"] += v; tot += v
for a in agg.values():
    a["share"] = a["us"] / tot
json.dump({"total_us": tot, "kernels": collections.OrderedDict(sorted(agg.items(), key=lambda kv: -kv[1]["us"]))},
          open(os.path.join(P, f"{tag}_launch_summary.json"), "w"), indent=1)
print("estep under ncu: %.1f us, %.1f MB DRAM" % (tot_us, tot_b / 1e6))
for n, a in per.items():
    print(f"  {n:40s} x{a['launches']:2d} {a['us']:8.1f} us {100 * a['share_of_estep']:5.1f} %  {a['dram_bytes'] / 1e6:8.1f} MB")
