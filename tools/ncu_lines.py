"""Aggregate an ncu report's per-instruction samples by CUDA source line (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 18
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print("=====", rows[1][1] if len(rows) > 1 else "?")
hdr = rows[2]; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
agg = []
for r in rows[3:]:
    if len(r) < iI + 1 or r[0] in ("File Path", "Function Name", "Line No") or r[2] != "-": continue
    try: agg.append((int(r[iS]), int(r[iI]), r[0], r[1][:100]))
    except ValueError: pass
ts = sum(a[0] for a in agg); ti = sum(a[1] for a in agg)
print("total samples", ts, "total inst", ti)
for a in sorted(agg, reverse=True)[:top]:
    print(f"{a[0]/ts*100:5.1f}% smp {a[1]/ti*100:5.1f}% inst  L{a[2]:>4}: {a[3]}")
