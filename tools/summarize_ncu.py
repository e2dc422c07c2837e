"""Summarise an `ncu --set full` report (read with `ncu -i X.ncu-rep --page raw --csv`) into the handful of metrics the
roofline discussion uses.  usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/xyz.txt"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks/SM"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors (32B)"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA load bytes"),
    ("smsp__inst_executed_op_tma_ld.sum", "TMA load instructions"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % of elapsed"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor (hmma subpipe) inst % of peak while active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__cycles_active.avg", "SMSP active cycles"), ("sm__cycles_elapsed.max", "elapsed cycles"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall samples: long scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "stall samples: barrier"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall samples: wait"),
    ("smsp__pcsamp_warps_issue_stalled_membar", "stall samples: membar"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall samples: short scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "stall samples: selected (issuing)"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("# %s — %d kernel launches captured with `ncu --set full --clock-control none`" % (path, len(rows) - 2))
    for r in rows[2:]:
        print("\n## %s   grid %s" % (r[col["Kernel Name"]][:110], r[col.get("Grid Size", 0)]))
        for key, label in KEYS:
            hits = [h for h in hdr if h == key or h.endswith("." + key)]
            if hits:
                i = col[hits[0]]
                print("  %-55s %s %s" % (label, r[i], units[i]))


if __name__ == "__main__":
    main(sys.argv[1])
