#!/usr/bin/env python3
"""Extracts the judged subset of metrics from an `ncu --set full` report into a small CSV.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/<name>.csv [launch index]
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_requests_srcunit_tex_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_red.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


def facts(rep, workload, tier, out_json, source, rays=None):
    """Adds the per-launch counters bench.py quotes (DRAM traffic, issue-slot and pipe utilisation) to
    profiles/ncu_facts.json under `workload`."""
    import json
    import os

    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    d = dict(zip(rows[0], rows[2]))
    u = dict(zip(rows[0], rows[1]))

    def bytes_of(k):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
        return float(d[k]) * scale

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "flatmatch-global-illumination_b200"))
    import fmgi

    entry = {
        "kernel": d["Kernel Name"], "tier": tier, "source": source,
        # the build the capture belongs to: must be run right after the capture, on the same tree
        "src_hash": fmgi.lib().fmgi_source_hash().decode(),
        "pipe_fmaheavy_pct": (float(d["sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active"]) if "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active" in d else None),
        "pipe_xu_pct": float(d.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "nan")),
        "warps_eligible_per_cycle": float(d.get("smsp__warps_eligible.avg.per_cycle_active", "nan")),
        "duration_ms": float(d["gpu__time_duration.sum"]) * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[u["gpu__time_duration.sum"]],
        "dram_bytes_per_launch": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
        "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "pipe_alu_pct": float(d["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]),
        "pipe_fma_pct": float(d["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]),
        "pipe_lsu_pct": float(d["sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]),
        "smem_wavefronts_pct": float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]),
        "threads_per_instruction": float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]),
        "registers_per_thread": int(float(d["launch__registers_per_thread"])),
        "l1_data_pipe_pct": float(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]),
        "l1_load_hit_pct": float(d["l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct"]),
        "warp_inst_per_launch": float(d["smsp__inst_executed.sum"]),
    }
    if rays:       # rays traced by the captured launch (bench.py: rays_per_s x kernel_ms_per_step / passes)
        entry["rays_per_launch"] = rays
        entry["warp_inst_per_ray"] = entry["warp_inst_per_launch"] / rays
    allf = json.load(open(out_json)) if os.path.exists(out_json) else {}
    allf[workload] = entry
    json.dump(allf, open(out_json, "w"), indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1))


def main():
    if sys.argv[1] == "--facts":            # --facts <rep> <workload> <tier> <source label> [rays in the launch]
        return facts(sys.argv[2], sys.argv[3], sys.argv[4], "profiles/ncu_facts.json", sys.argv[5],
                     float(sys.argv[6]) if len(sys.argv) > 6 else None)
    rep, out = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2 + launch]
    with open(out, "w") as f:
        f.write("metric,unit,value\n")
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or h == "Kernel Name":
                f.write(f"{h},{u},{v}\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
