"""Summarise an ncu report (ncu -i X.ncu-rep --page raw --csv): one block per kernel launch with the counters DESIGN.md / bench.py quote.
    python scripts/ncu_summary2.py report.ncu-rep [out.txt] [traffic.json positions]
"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
ix = {n: i for i, n in enumerate(h)}
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_utchmma.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
out = []
traffic = {}
for r in rows[2:]:
    if len(r) < len(h) - 2:
        continue
    name = r[ix["Kernel Name"]]
    out.append(f"== launch {r[ix['ID']]}: {name[:110]}")
    units = rows[1]
    for k in want:
        if k in ix and r[ix[k]] != "":
            out.append(f"   {k:92s} {r[ix[k]]} {units[ix[k]]}")
    try:
        b = float(r[ix["dram__bytes_read.sum"]].replace(",", "")) + float(r[ix["dram__bytes_write.sum"]].replace(",", ""))
        u = units[ix["dram__bytes_read.sum"]]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        traffic.setdefault(name, []).append(b * mult)
    except Exception:
        pass
text = "\n".join(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
else:
    print(text)
if len(sys.argv) > 4:
    mg = sum(v[0] for k, v in traffic.items() if "k_movegen" in k)
    ev = sum(sum(v) for k, v in traffic.items() if "k_eval_tc" in k)
    json.dump({"positions": int(sys.argv[4]), "movegen_bytes": mg, "eval_bytes": ev, "source": rep.split("/")[-1] + " (ncu --set full, scripts/prof_movegen21.py, one fused step; dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over the kernels of the step)",
               "per_kernel": {k[:90]: v for k, v in traffic.items()}}, open(sys.argv[3], "w"), indent=1)
