"""Print the key metrics of every kernel in an .ncu-rep (read here, on the CPU box):
    python scripts/ncu_raw_summary.py gpurun_out/prof.ncu-rep [frames_per_launch]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
M = [('gpu__time_duration.sum', 'time'), ('sm__inst_executed.sum', 'warp instr'), ('sm__inst_executed.avg.per_cycle_elapsed', 'IPC/SM'),
     ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM thr %'), ('smsp__issue_active.avg.pct', 'issue active %'),
     ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'), ('launch__registers_per_thread', 'regs'),
     ('launch__occupancy_limit_registers', 'occ lim regs'), ('launch__occupancy_limit_shared_mem', 'occ lim smem'),
     ('dram__bytes_read.sum', 'dram rd'), ('dram__bytes_write.sum', 'dram wr'), ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
     ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem wavefronts'), ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
     ('sm__inst_executed_pipe_alu.sum', 'pipe alu'), ('sm__inst_executed_pipe_fma.sum', 'pipe fma'), ('sm__inst_executed_pipe_lsu.sum', 'pipe lsu'),
     ('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'alu pipe %'), ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma pipe %'),
     ('l1tex__lsu_writeback_active_mem_lg.sum', ''), ('sm__cycles_elapsed.max', 'cycles')]
for r in rows[2:]:
    print('##', r[idx['Kernel Name']][:90], '| grid', r[idx['launch__grid_size']], 'block', r[idx['launch__block_size']])
    for m, label in M:
        if m in idx and label:
            v = r[idx[m]]
            extra = ''
            if frames and m in ('sm__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum'):
                try: extra = f'  ({float(v.replace(",", ""))/frames:.0f} per frame)'
                except ValueError: pass
            print(f'   {label:22s} {v} {units[idx[m]]}{extra}')
