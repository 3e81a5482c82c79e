#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a small JSON/markdown table: python tools/ncu_summary.py rep [out.json]"""
import csv, json, subprocess, sys
WANT = {
 "gpu__time_duration.sum": "time_us",
 "dram__bytes_read.sum": "dram_read",
 "dram__bytes_write.sum": "dram_write",
 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
 "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
 "lts__t_bytes.sum": "l2_bytes",
 "lts__t_sectors_srcunit_tex_op_read.sum": "l2_tex_read_sectors",
 "lts__t_sectors_op_read.sum": "l2_read_sectors",
 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct_active",
 "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_pct",
 "sm__inst_executed_pipe_uniform.sum": "uniform_inst",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
 "sm__cycles_active.avg": "sm_cycles_active",
 "sm__cycles_elapsed.max": "sm_cycles_elapsed",
 "launch__grid_size": "grid",
 "launch__registers_per_thread": "regs",
 "launch__shared_mem_per_block_dynamic": "dyn_smem",
 "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
 "smsp__cycles_active.avg": "smsp_cycles_active",
 "gpc__cycles_elapsed.max": "gpc_cycles",
 "sm__cycles_elapsed.avg.per_second": "sm_hz",
 "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
 "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
 "smsp__inst_executed.sum": "warp_instructions",
}
def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:60]}
        for i, h in enumerate(hdr):
            if h in WANT:
                d[WANT[h]] = f"{r[i]} {units[i]}".strip()
        res.append(d)
    if len(sys.argv) > 2:
        json.dump(res, open(sys.argv[2], "w"), indent=1)
    for d in res:
        print(json.dumps(d))
    if "--list" in sys.argv:
        for h in hdr: print(h)
main()
