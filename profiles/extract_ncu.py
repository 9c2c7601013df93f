"""Key metrics of a `ncu --set full` report -> small CSV (the .ncu-rep files themselves stay in gpurun_out/, untracked).
Usage: python profiles/extract_ncu.py report.ncu-rep > profiles/r01_ncu_full_<name>.csv"""
import csv
import subprocess
import sys

NAMES = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
         "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
         "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
         "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
         "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.sum",
         "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.per_cycle_elapsed",
         "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.peak_sustained",
         "TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg",
         "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_write_bytes.sum",
         "l1tex__m_xbar2l1tex_read_sectors.avg.pct_of_peak_sustained_elapsed",
         "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
         "smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
         "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
for n in NAMES:
    if n in hdr:
        i = hdr.index(n)
        w.writerow([n, units[i]] + [r[i][:120] for r in data])
