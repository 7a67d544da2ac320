"""Condensed per-launch table from an `ncu --page raw --csv` export:
python tools/ncu_table.py gpurun_out/<tag>_qg_raw.csv > profiles/<tag>_qg_ncu_table.txt"""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("sm__inst_executed.avg.per_cycle_elapsed", "ipc")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
use = [(c, n) for c, n in COLS if c in col]
print("# " + sys.argv[1] + "  (ncu --set full --clock-control none; cold-cache, serialised launches)")
print("kernel | " + " | ".join(f"{n} [{units[col[c]]}]" if units[col[c]] else n for c, n in use))
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    print(name + " | " + " | ".join(r[col[c]] for c, n in use))
