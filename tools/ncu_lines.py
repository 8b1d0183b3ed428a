"""Per-source-line stall samples from an .ncu-rep captured with --import-source on (read here, no GPU needed):
python tools/ncu_lines.py rep.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
cur_file = "?"
lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] == "":
        continue  # SASS rows
    d = dict(zip(hdr, r))
    try:
        samples = int(d["# Samples"])
    except ValueError:
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    lines.append((samples, cur_file, r[0], r[1].strip()[:90], stalls, d.get("Instructions Executed", "0"), d.get("L1 Wavefronts Shared Excessive", "0")))
tot = sum(x[0] for x in lines)
print("total samples", tot)
for s, f, ln, src, st, ie, exc in sorted(lines, reverse=True)[:top]:
    tops = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% %s:%s  inst=%s excess_smem=%s  %s\n        %s" % (100.0 * s / tot, f, ln, ie, exc, " ".join("%s=%d" % kv for kv in tops), src))
if len(sys.argv) > 3:
    # line-range totals: "name:lo-hi,name:lo-hi" for the main file (argv[4] or sy2sb.cu)
    mainf = sys.argv[4] if len(sys.argv) > 4 else "sy2sb.cu"
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        s = sum(x[0] for x in lines if x[1] == mainf and lo <= int(x[2]) <= hi)
        print("%-12s %5.1f%%" % (name, 100.0 * s / tot))
    print("other files %5.1f%%" % (100.0 * sum(x[0] for x in lines if x[1] != mainf) / tot))
