"""Text table of the key ncu metrics of every launch in an .ncu-rep (read here, no GPU):
usage: python scripts/ncu_kernel_table.py gpurun_out/x.ncu-rep [names of the launches ...] > profiles/x.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
labels = sys.argv[2:]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
COLS = [('gpu__time_duration.sum', 'time', 1), ('launch__grid_size', 'grid', 1), ('launch__registers_per_thread', 'regs', 1),
        ('launch__occupancy_limit_shared_mem', 'occ_smem', 1), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_act%', 1),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_thru%', 1), ('smsp__inst_executed.sum', 'warp_inst', 1),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%', 1), ('dram__bytes_read.sum', 'dram_rd', 1),
        ('dram__bytes_write.sum', 'dram_wr', 1), ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 1),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_conflicts', 1), ('lts__t_sector_hit_rate.pct', 'l2_hit%', 1)]
cols = [(k, n, f) for k, n, f in COLS if k in idx]
print(f'# {rep}: ncu --set full --clock-control none (per launch; batch of 64 frames, Lite0)')
print(f'{"launch":18s} ' + ' '.join(f'{n:>14s}' for _, n, _ in cols))
print(f'{"(unit)":18s} ' + ' '.join(f'{units[idx[k]][:14]:>14s}' for k, _, _ in cols))
for i, r in enumerate(data):
    def val(k, f):
        try:
            return float(r[idx[k]].replace(',', '')) * f
        except ValueError:
            return float('nan')
    name = labels[i] if i < len(labels) else r[idx['Kernel Name']][:18]
    print(f'{name:18s} ' + ' '.join(f'{val(k, f):14.2f}' for k, _, f in cols))
