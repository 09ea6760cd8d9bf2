"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into one line per captured launch.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[idx["Kernel Name"]][:48]
    parts = [name]
    for k in KEYS:
        if k in idx:
            parts.append(f"{k.split('.')[0]}={r[idx[k]]}{units[idx[k]]}")
    # top stall reasons of the sampled warps (pc sampling, all samples)
    st = []
    for h_ in hdr:
        if h_.startswith("smsp__pcsamp_warps_issue_stalled_") and not h_.endswith("_not_issued"):
            try:
                st.append((float(r[idx[h_]].replace(",", "")), h_[len("smsp__pcsamp_warps_issue_stalled_"):]))
            except ValueError:
                pass
    tot = sum(v for v, _ in st)
    if tot > 0:
        st.sort(reverse=True)
        parts.append("stalls: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in st[:4]))
    print(" | ".join(parts))
