"""Summarise ncu reports (--set full) into a markdown table + a small JSON, for profiles/.
usage: python tools/profile_summary.py out_prefix report1.ncu-rep [report2 ...]"""
import csv, io, json, subprocess, sys

WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("smsp__inst_executed.sum", "warp_inst")]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[H.index("Kernel Name")].split("(")[0].replace("void ", "").replace("mcs::", "")}
        for m, k in WANT:
            if m in H:
                i = H.index(m)
                d[k] = r[i]
                d[k + "_unit"] = U[i]
        res.append(d)
    return res


def tobytes(v, unit):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    prefix, reps = sys.argv[1], sys.argv[2:]
    allk = []
    for rep in reps:
        for d in load(rep):
            d["report"] = rep.split("/")[-1]
            allk.append(d)
    lines = ["| kernel | grid | regs | time | DRAM rd+wr (MB) | DRAM % | issue % | warps % | ALU % | FP64 % | XU % | L1TEX % | warp-inst |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    js = []
    for d in allk:
        mb = (tobytes(d.get("dram_rd", 0), d.get("dram_rd_unit", "byte")) + tobytes(d.get("dram_wr", 0), d.get("dram_wr_unit", "byte"))) / 1e6
        t = f'{float(d["time"]):.1f} {d["time_unit"]}'
        lines.append(f'| {d["kernel"]} | {d.get("grid","")} | {d.get("regs","")} | {t} | {mb:.1f} | {float(d.get("dram%",0)):.1f} | '
                     f'{float(d.get("issue%",0)):.1f} | {float(d.get("warps%",0)):.1f} | {float(d.get("alu%",0)):.1f} | {float(d.get("fp64%",0)):.1f} | '
                     f'{float(d.get("xu%",0)):.1f} | {float(d.get("l1tex%",0)):.1f} | {float(d.get("warp_inst",0))/1e6:.1f} M |')
        js.append({"kernel": d["kernel"], "grid": d.get("grid"), "time": t, "dram_bytes": mb * 1e6, "issue_pct": float(d.get("issue%", 0)), "report": d["report"]})
    open(prefix + ".md", "w").write("\n".join(lines) + "\n")
    json.dump(js, open(prefix + ".json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
