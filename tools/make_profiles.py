"""Turn the raw ncu outputs of a round into the committed summaries under profiles/.
usage: python tools/make_profiles.py <launches.csv> <rows_full.ncu-rep> <round tag, e.g. r01>"""
import csv, json, subprocess, sys

launch_csv, rep, tag = sys.argv[1:4]
# ---- launch list -------------------------------------------------------------------------------------------------
lines = [l for l in open(launch_csv) if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
    seq.append((row["Kernel Name"], ms))
agg = {}
for n, ms in seq:
    k = n.split("(")[0][:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(v[1] for v in agg.values())
out = [f"# {tag} — ncu launch list of `python bench.py --steps 2 --warmup 1 --cpu-sample none` (C3, 1x B200)", "",
       f"`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, not absolutes). Raw: {tag}_launches_bench_c3.csv", "",
       "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {v[0]} | {v[1]:.3f} | {v[1] / tot * 100:.1f}% |")
out.append(f"| **all** | {len(seq)} | {tot:.3f} | 100% |")
share = sum(v[1] for k, v in agg.items() if "k_rows" in k) / tot
out += ["", f"`k_rows` (fused A'^T B' count + LLR + top-k, all work bins) = **{share * 100:.1f}%** of the device time under ncu; the live CUDA-event "
        f"bracket of bench.py (`config.stage_ms_last_resident_step`) gives row_kernels / step = the same share within a few points "
        "(the bins run concurrently on separate streams there, serialised under ncu)."]
open(f"profiles/{tag}_launches_c3.md", "w").write("\n".join(out) + "\n")

# ---- ncu --set full of the row kernels --------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr = r[0]
col = lambda name: hdr.index(name)
md = [f"# {tag} — `ncu --set full` of the fused row kernel `k_rows` (C3, 1x B200; one launch per work bin of one indicator)", "",
      "Command: `ncu --set full --clock-control none --import-source on -k regex:k_rows python bench.py --steps 1 --warmup 1 --cpu-sample none` "
      "(the .ncu-rep stays in gpurun_out/, not in git).", "",
      "| launch | grid x block | ms | DRAM rd MB | DRAM wr MB | warps active % | issue active % | fp64 pipe % | warp-inst M | regs | stalls: barrier / wait / long_sb / short_sb / branch |",
      "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
tr = tw = tt = 0.0
n = 0
for row in r[2:]:
    name = row[col("Kernel Name")]
    if "k_rows" not in name:
        continue
    t = float(row[col("gpu__time_duration.sum")])
    tu = r[1][col("gpu__time_duration.sum")]
    t = t / 1e6 if tu.startswith("n") else (t / 1e3 if tu.startswith("u") else t)   # -> ms
    if t < 0.01:
        continue  # empty bin
    n += 1
    def g(m):
        v = float(row[col(m)] or 0)
        u = r[1][col(m)]
        return v / 1e6 if u == "byte" else (v / 1e3 if u == "Kbyte" else (v * 1e3 if u == "Gbyte" else v))   # bytes -> MB
    s = lambda m: row[col("smsp__pcsamp_warps_issue_stalled_" + m)]
    md.append(f"| {name.replace('void ', '').replace('(RowArgs)', '')} | {row[col('Grid Size')]} x {row[col('Block Size')]} | {t:.3f} | {g('dram__bytes_read.sum'):.1f} | "
              f"{g('dram__bytes_write.sum'):.1f} | {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | {g('smsp__inst_executed.sum') / 1e6:.0f} | {row[col('launch__registers_per_thread')]} | "
              f"{s('barrier')} / {s('wait')} / {s('long_scoreboard')} / {s('short_scoreboard')} / {s('branch_resolving')} |")
    tr += g("dram__bytes_read.sum"); tw += g("dram__bytes_write.sum"); tt += t
ind = max(n // 8, 1) if n > 8 else 1
md += ["", f"Captured launches: {n} (= {ind} indicator(s)); {tt:.2f} ms serialised; DRAM traffic {tr + tw:.0f} MB (read {tr:.0f}, write {tw:.0f}).",
       "SURVEY 8(d) algorithmic bytes of one C3 indicator are ~664 MB: the gathers of B' (20 MB after downsampling) and of the per-column terms hit "
       "the 126 MB L2, so the kernel moves LESS than its algorithmic bytes from HBM.  What binds it (DESIGN.md 3.2): CTA-owned bins wait at the "
       "barrier that ends the count phase (`barrier` column), warp-owned bins are limited by shared memory per row (warps active) and spend ~30 % "
       "of their instructions in the final sort; fp64 pipe utilisation is low after the level-1 integer cut (12.7 % of the cells reach the LLR)."]
open(f"profiles/{tag}_k_rows_ncu_full.md", "w").write("\n".join(md) + "\n")
import hashlib, os
build = hashlib.sha1(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "universal_recommender_b200", "csrc",
                                       "cco_kernels.cuh"), "rb").read()).hexdigest()[:12]
json.dump({"source": f"profiles/{tag}_k_rows_ncu_full.md", "workload": "C3", "dram_bytes_per_indicator": (tr + tw) * 1e6 / ind,
           "launches_per_indicator": n / ind, "build": build}, open(f"profiles/{tag}_k_rows_traffic.json", "w"))
print("\n".join(md[6:]))
