"""Aggregate an ncu report's per-line executed instructions / stall samples (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [launch_index] [top_n]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; skip = sys.argv[2] if len(sys.argv) > 2 else "0"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict(); fname = ""
H = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 3 and r[0] == "Line No":
        H = r; ii = H.index("Instructions Executed"); ws = H.index("Warp Stall Sampling (All Samples)"); ti = H.index("Thread Instructions Executed"); continue
    if H and len(r) > ii and r[0].isdigit() and r[2] == "-":      # per-source-line summary rows
        key = (fname, int(r[0]), r[1].strip())
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += int(r[ii] or 0); a[1] += int(r[ws] or 0); a[2] += int(r[ti] or 0)
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print(f"total warp-instructions {tot}, stall samples {tots}")
for (f, ln, src), (n, w, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n / tot * 100:5.1f}% inst {w / max(tots, 1) * 100:5.1f}% stall  thr/inst {t / max(n, 1):4.1f}  {f}:{ln:<4d} {src[:105]}")
