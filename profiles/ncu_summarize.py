"""Summary of an `ncu --set full` capture: per launch duration, grid, registers, occupancy, issue rate, DRAM bytes and
tensor-pipe activity -> JSON; plus the DRAM traffic of the warp QP kernels of one working-set round (the `traffic`
of the bench line's roofline entry).

    ncu -i gpurun_out/<rep>.ncu-rep --page raw --csv > profiles/ncu_raw_r02_s3.csv
    python profiles/ncu_summarize.py profiles/ncu_raw_r02_s3.csv profiles/ncu_summary_r02_s3.json profiles/traffic_r02.json
"""
import csv
import json
import re
import sys

src, out_summary, out_traffic = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
h, units = rows[0], rows[1]
ki = h.index("Kernel Name")
want = {"us": "gpu__time_duration.sum", "grid": "launch__grid_size", "regs": "launch__registers_per_thread",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "issue_active_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "dram_read_mb": "dram__bytes_read.sum", "dram_write_mb": "dram__bytes_write.sum",
        "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dmma_inst_pct": "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "l2_hit_pct": "lts__t_sector_hit_rate.pct",
        "stall_long_scoreboard_per_issue": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "stall_wait_per_issue": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "stall_barrier_per_issue": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "stall_no_instruction_per_issue": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def mb(col, r):
    v, u = num(r[h.index(col)]), units[h.index(col)]
    if v is None:
        return None
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)


summary, order = {}, []
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("revs::", "").replace("<unnamed>::", "").strip()
    name = re.sub(r"\(int\)|\(bool\)", "", name)
    e = {}
    for k, col in want.items():
        if col in h:
            e[k] = mb(col, r) if k.startswith("dram_") and k.endswith("_mb") else num(r[h.index(col)])
    summary.setdefault(name, []).append(e)
    order.append(name)
json.dump({"source": src, "launch_order": order, "kernels": summary}, open(out_summary, "w"), indent=1)

# one working-set round = the first run of consecutive warp-kernel launches
warp = [n for n in order if n.startswith("utility_qp_warp_kernel") or n.startswith("utility_qp_fast_kernel")]
first_round, seen = {}, set()
for n in order:
    if (n.startswith("utility_qp_warp_kernel") or n.startswith("utility_qp_fast_kernel")) and n not in seen:
        seen.add(n)
        e = summary[n][0]
        first_round[n] = {"dram_read_MB": round(e["dram_read_mb"], 3), "dram_write_MB": round(e["dram_write_mb"], 3), "us": e["us"]}
tr = {"source": src + " (ncu --set full --clock-control none, single pipeline, first working-set round of one ADMM iteration)"}
tr.update(first_round)
tr["utility_qp_warp_kernel_bytes_per_round"] = int(1e6 * sum(v["dram_read_MB"] + v["dram_write_MB"] for v in first_round.values()))
tr["utility_qp_warp_kernel_bytes_per_round_note"] = ("warp QP kernels of one round over ALL zones of the GPU (single pipeline); with K pipelines a launch "
                                                      "group covers 1/K of the zones")
for n in ("home_solve_kernel<3>", "dual_update_kernel<3, 1>", "qp_init_kernel", "screen_tc5_kernel"):
    if n in summary:
        e = max(summary[n], key=lambda x: x["us"])
        tr[n] = {"dram_read_MB": round(e["dram_read_mb"], 3), "dram_write_MB": round(e["dram_write_mb"], 3), "us": e["us"]}
json.dump(tr, open(out_traffic, "w"), indent=1)
print(json.dumps(tr, indent=1))
