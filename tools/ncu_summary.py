"""Summarise `ncu --set full` reports (tools/ncu_kernels.sh) into the per-kernel text files committed under profiles/:
duration, DRAM bytes, L2 -> SM and SM -> L2 bytes, tensor-pipe activity, issue activity, top warp stall reasons.
usage: python tools/ncu_summary.py <report.ncu-rep> [...]"""
import csv
import re
import subprocess
import sys

KEEP = [r"^gpu__time_duration\.sum$", r"^sm__cycles_elapsed\.avg$", r"^sm__cycles_active\.avg$",
        r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^l1tex__m_xbar2l1tex_read_bytes\.sum$", r"^l1tex__m_l1tex2xbar_write_bytes\.sum$", r"^lts__t_sector_hit_rate\.pct$",
        r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active$",
        r"^sm__pipe_tensor_subpipe_hmma_cycles_active\.avg\.pct_of_peak_sustained_active$",
        r"^sm__inst_executed_pipe_tmem\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
        r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^launch__registers_per_thread$",
        r"^launch__grid_size$", r"^launch__block_size$", r"^launch__cluster_size$", r"^launch__shared_mem_per_block_dynamic$",
        r"^smsp__sass_inst_executed_op_tmem_ldt\.sum$", r"^smsp__sass_inst_executed_op_tmem_stt\.sum$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$"]
STALL = re.compile(r"^smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio$")


def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        name_col = hdr.index("Kernel Name")
        print("# %s" % path.split("/")[-1])
        for r in rows[2:]:
            print("## launch id %s  %s" % (r[hdr.index("ID")], r[name_col][:110]))
            stalls = []
            for i, h in enumerate(hdr):
                if any(re.search(p, h) for p in KEEP):
                    print("  %-82s %16s %s" % (h, r[i], units[i]))
                m = STALL.match(h)
                if m:
                    try:
                        stalls.append((float(r[i]), m.group(1)))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("  top stall reasons (warps stalled per issue-active cycle): " +
                  ", ".join("%s %.2f" % (n, v) for v, n in stalls[:6]))
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
