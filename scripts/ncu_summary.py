"""Summarise an `ncu --set full` report: one block of key metrics per captured kernel.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % of elapsed"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % of active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_bytes.sum.per_second", "L2 byte rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) CTAs"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) CTAs"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long scoreboard %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# %s : %d kernel launch(es), `ncu --set full --clock-control none`" % (rep.split("/")[-1], len(data)))
    for r in data:
        print("\n## %s  grid %s block %s" % (r[ix["Kernel Name"]][:150], r[ix["Grid Size"]], r[ix["Block Size"]]))
        for k, label in KEYS:
            if k in ix:
                print("  %-36s %14s %s" % (label, r[ix[k]], units[ix[k]]))


if __name__ == "__main__":
    main()
