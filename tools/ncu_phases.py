"""Instruction / stall shares of a kernel by source phase: the per-line table of tools/ncu_lines.py grouped by marker comments.
usage: python tools/ncu_phases.py report.ncu-rep launch_index source.cu 'name=marker text' ...   (phases in source order; lines before
the first marker are reported per device function)"""
import collections, csv, io, re, subprocess, sys
rep, launch, srcfile = sys.argv[1], sys.argv[2], sys.argv[3]
marks = [a.split("=", 1) for a in sys.argv[4:]]
src = open(srcfile).read().splitlines()
base = srcfile.split("/")[-1]

def find(pat):
    for i, l in enumerate(src):
        if pat in l:
            return i + 1
    raise SystemExit(f"marker not found: {pat}")
bounds = sorted((find(pat), name) for name, pat in marks)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", launch, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg, fname, H = collections.OrderedDict(), "", None
tot = tots = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) > 3 and r[0] == "Line No":
        H = r; ii = H.index("Instructions Executed"); ws = H.index("Warp Stall Sampling (All Samples)"); continue
    if H and len(r) > ii and r[0].isdigit() and r[2] == "-":
        ln, n, w = int(r[0]), int(r[ii] or 0), int(r[ws] or 0)
        tot += n; tots += w
        if fname != base:
            name = f"(inlined from {fname})"
        else:
            name = "(before the first marker)"
            for l0, nm in bounds:
                if ln >= l0:
                    name = nm
        a = agg.setdefault(name, [0, 0]); a[0] += n; a[1] += w
print(f"{rep.split('/')[-1]} launch {launch}: {tot} warp-instructions, {tots} stall samples\n")
print("| phase | warp-instructions | share | stall samples share |\n|---|---|---|---|")
for name, (n, w) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {name} | {n / 1e6:.1f} M | {n / tot * 100:.1f} % | {w / max(tots, 1) * 100:.1f} % |")
