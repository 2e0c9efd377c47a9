"""Summarise an `ncu --set full` report (.ncu-rep) into the two committed files
bench.py and the judge read:
  profiles/<tag>_ncu_full_summary.csv   one row per captured kernel (duration, DRAM
                                        bytes, DRAM %, issue %, registers, stalls)
  profiles/r2_ncu_traffic.json          {"kernels": {bench kernel id: {"dram_bytes": ...}}}
Usage (build container, no GPU needed):
  python profiles/ncu_traffic.py gpurun_out/<rep>.ncu-rep [...] --tag r2
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
# bench.py kernel ids (sb_profile slots) by kernel-name substring
IDS = [("thth_eig", "thth_eig"), ("thth_build", "thth_build"), ("row_fft_r2c", "cs_rows"),
       ("ColAStore", "cs_colA"), ("CsStore", "cs_colB")]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tag = "r2"
    if "--tag" in sys.argv:
        tag = sys.argv[sys.argv.index("--tag") + 1]
        args = [a for a in args if a != tag]
    out_rows, traffic = [], {}
    for rep in args:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True,
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            out_rows.append([name] + [r[col[k]] if k in col else "" for k in KEEP])
            def val(k):
                return float(r[col[k]]) * UNIT.get(units[col[k]], 1.0)
            for sub, kid in IDS:
                if sub in name:
                    traffic[kid] = {"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                                    "duration_ms_under_ncu": float(r[col["gpu__time_duration.sum"]]),
                                    "kernel": name.split("(")[0], "report": os.path.basename(rep)}
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "%s_ncu_full_summary.csv" % tag), "w", newline="") as fh:
        wr = csv.writer(fh)
        wr.writerow(["Kernel Name"] + KEEP)
        wr.writerows(out_rows)
    with open(os.path.join(here, "r2_ncu_traffic.json"), "w") as fh:
        json.dump({"source": "ncu --set full --clock-control none of `python bench.py --steps 1 "
                             "--warmup 2 --no-cpu --no-strong --no-extra`, one launch per kernel "
                             "(%s)" % ", ".join(os.path.basename(a) for a in args),
                   "kernels": traffic}, fh, indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
