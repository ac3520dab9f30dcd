"""Key metrics per kernel launch of an ncu report -> JSON lines (read on the CPU box).
usage: python profiles/extract_ncu.py report.ncu-rep > summary.jsonl"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active": "dmma_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__inst_executed.sum": "warp_insts",
}
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(io.StringIO(out)))
hdr, units = rd[0], rd[1]
for row in rd[2:]:
    d = dict(zip(hdr, row))
    u = dict(zip(hdr, units))
    rec = {"kernel": d["Kernel Name"].split("(")[0][:70]}
    for k, name in KEYS.items():
        if k in d and d[k] not in ("", "n/a"):
            v = float(d[k].replace(",", ""))
            if name == "time_us":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u[k], 1.0)
            if name.endswith("_MB"):
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u[k], 1.0)
            rec[name] = round(v, 3)
    st = {h.split("stalled_")[1]: float(v.replace(",", "")) for h, v in d.items()
          if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and v not in ("", "n/a")}
    tot = sum(st.values()) or 1.0
    rec["stalls_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:6]}
    print(json.dumps(rec))
