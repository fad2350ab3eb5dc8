"""Per-kernel summary of an ncu report -> JSON (and optionally profiles/traffic.json).

    python tools/ncu_summary.py <report.ncu-rep> <out.json> [--traffic profiles/traffic.json]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    res, traffic = [], {}
    for r in rows[2:]:
        d = {"kernel": r[ix["Kernel Name"]]}
        for k in KEYS:
            if k in ix:
                d[k] = r[ix[k]]
                if k.startswith("dram__bytes"):
                    d[k + ".unit"] = units[ix[k]]
        res.append(d)
        b = sum(float(r[ix[k]]) * UNIT_SCALE.get(units[ix[k]], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        name = d["kernel"]
        for tag in ("rows_fwd", "rows_inv", "x_c2r", "x_r2c", "cols"):
            if tag in name:
                traffic.setdefault(tag, b)
    json.dump(res, open(out, "w"), indent=1)
    if "--traffic" in sys.argv:
        json.dump(traffic, open(sys.argv[sys.argv.index("--traffic") + 1], "w"), indent=1)
    for d in res:
        print(d["kernel"][:60], d.get("gpu__time_duration.sum"))


if __name__ == "__main__":
    main()
