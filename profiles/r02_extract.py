"""Summaries of the round-2 `ncu --set full` reports (gpurun_out/r02_*.ncu-rep, captured by profiles/r02_capture.sh):
    python profiles/r02_extract.py gpurun_out/r02_seq.ncu-rep [...] > profiles/r02_ncu_summary.md
One block per captured launch with the metrics the profiling recipe names (B200_PROFILING.md)."""
import csv
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'duration'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__registers_per_thread', 'regs/thread'), ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
        ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
        ('l1tex__t_sector_hit_rate.pct', 'L1 hit rate %'),
        ('smsp__inst_executed.sum', 'instructions'),
        ('sm__inst_executed_pipe_tensor.sum', 'tensor instructions')]

for rep in sys.argv[1:]:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        print('## %s: empty report\n' % rep)
        continue
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print('## %s\n' % rep.split('/')[-1])
    for r in rows[2:]:
        name = r[col['Kernel Name']]
        print('* `%s`' % name[:140])
        parts = []
        for key, label in WANT:
            if key in col and r[col[key]] != '':
                parts.append('%s %s%s' % (label, r[col[key]], (' ' + units[col[key]]) if units[col[key]] not in ('', '%') else ''))
        print('  ' + '; '.join(parts))
    print()
