#!/usr/bin/env python
"""ncu -i <rep> --page raw --csv  ->  profiles/<name>.summary.csv (metric, unit, value), keeping the metrics that matter
for this kernel: DRAM bytes / throughput, L2 hit rate, L2->SM bytes (xbar2l1tex, the TMA loads), tensor-pipe activity, issue stalls, launch shape.
Usage (here, no GPU needed):  python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_x.summary.csv"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(r"^(dram__bytes|dram__cycles_elapsed|dram__throughput|gpu__dram_throughput|gpu__time_duration|l1tex__throughput|"
                  r"l1tex__data_pipe_lsu_wavefronts(_mem_shared)?\.sum$|launch__|lts__t_sector_hit_rate|lts__t_bytes|"
                  r"lts__t_sectors_srcunit_tex_op_read|lts__throughput|lts__cycles_elapsed|sm__cycles_(active|elapsed)|"
                  r"sm__inst_executed_pipe_tensor|sm__pipe_tensor|sm__throughput|sm__warps_active|smsp__average_warp|"
                  r"smsp__inst_executed\.sum|smsp__cycles_active\.avg$|"
                  r"l1tex__m_xbar2l1tex_read_bytes(_mem_global_op_tma_ld)?\.sum(\.per_second)?$|lts__t_sectors_srcunit_tex\.sum$|"
                  r"lts__t_sectors\.sum$)")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for h, u, v in zip(hdr, units, vals):
            if h in ("Kernel Name", "Block Size", "Grid Size") or KEEP.match(h):
                w.writerow([h, u, v])
    print(out)


if __name__ == "__main__":
    main()
