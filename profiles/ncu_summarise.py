"""Summarise an `ncu --set full` report here on the CPU box (no GPU needed):

    python profiles/ncu_summarise.py gpurun_out/r2_step.ncu-rep > profiles/r2_ncu_kernels.json

One row per profiled launch: duration, DRAM bytes read / written and achieved GB/s, DRAM / L1 / L2 / tensor-pipe
utilisation, registers, achieved occupancy, issue-slot use and the top warp-stall reason (ncu's own metric names
are kept in the keys so the numbers can be found again in the report).
"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "smsp__cycles_active.avg": "smsp_cycles_active",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_hz",
    "l1tex__t_bytes.sum": "l1tex_bytes",
    "lts__t_bytes.sum": "l2_bytes",
}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}
    stall_cols = [h for h in head if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    if not stall_cols:
        stall_cols = [h for h in head if h.startswith("smsp__average_warp_latency_issue_stalled_")]
    res = []
    for r in rows[2:]:
        if len(r) < len(head):
            continue
        d = {"id": r[col["ID"]], "kernel": r[col["Kernel Name"]]}

        def num(name):
            try:
                return float(r[col[name]].replace(",", ""))
            except (KeyError, ValueError):
                return None
        for k, short in WANT.items():
            v = num(k)
            if v is not None:
                u = units[col[k]]
                if short == "duration":
                    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)   # -> us
                if short in ("dram_read", "dram_write", "l1tex_bytes", "l2_bytes"):
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
                    v = v * scale
                d[short] = v
        if d.get("duration") and d.get("dram_read") is not None:
            d["dram_gbs"] = (d["dram_read"] + d.get("dram_write", 0.0)) / (d["duration"] * 1e-6) / 1e9
        stalls = sorted(((num(c) or 0.0, c) for c in stall_cols), reverse=True)[:3]
        d["top_stalls"] = [{"reason": c.split("issue_stalled_")[1].split("_per_")[0].replace(".ratio", ""), "value": v} for v, c in stalls]
        res.append(d)
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
