"""ncu_stalls.py -- warp-state / memory-path summary of the kernels of an ncu report whose name matches a pattern:
stall reasons per issue, issue rate, occupancy, local- vs global-memory requests and their L1 hit rates, shared-memory wavefronts.
    python tools/ncu_stalls.py gpurun_out/r3_k3.ncu-rep describe [title]"""
import csv, io, re, subprocess, sys

rep, pat = sys.argv[1], re.compile(sys.argv[2])
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
units = dict(zip(rows[0], rows[1]))
want = [
    ("gpu__time_duration.sum", "time under ncu"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "CTAs / SM allowed by registers"),
    ("launch__occupancy_limit_shared_mem", "CTAs / SM allowed by shared memory"),
    ("launch__shared_mem_config_size", "shared-memory carve-out (KB)"),
    ("sm__warps_active.avg.per_cycle_active", "active warps / SM"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used (%)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (L1TEX: global / local memory)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (shared memory, MUFU)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "LOCAL (spill) load requests"),
    ("l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "LOCAL (spill) store requests"),
    ("l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "L1 hit rate, global loads (%)"),
    ("l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "L1 hit rate, local loads (%)"),
    ("smsp__sass_inst_executed_op_shared_ld.sum", "shared-memory load instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank-conflict wavefronts"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (%)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (%)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe (%)"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe (%)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (%)"),
]
name_col = hdr.index("Kernel Name")
out = [f"# {title}", ""]
for r in rows[2:]:
    if not r or not pat.search(r[name_col]):
        continue
    d = dict(zip(hdr, r))
    out += [f"## {d['Kernel Name'][:80]}  (launch id {d.get('ID', '?')}, grid {d.get('Grid Size', '?')})", "", "| metric | value |", "|---|---|"]
    for k, label in want:
        if k in d and d[k] != "":
            out.append(f"| {label} | {d[k]} {units.get(k, '')} |".replace("  |", " |"))
    out.append("")
print("\n".join(out))
