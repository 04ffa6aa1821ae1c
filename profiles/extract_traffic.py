"""Read an `ncu --set full` report (via `ncu -i X.ncu-rep --page raw --csv`) and write the per-launch DRAM traffic of the dominant
kernel (the rnn4 fused LSTM layer = rc_tc_kernel<...,1> launches with grid (40, M-tiles, 1)) to profiles/r01_tc_traffic.json."""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else 'profiles/r01_tc_traffic.json'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
sel = []
for r in rows[2:]:
    if 'rc_tc_kernel' in r[col['Kernel Name']] and ', 1>' in r[col['Kernel Name']] and r[col['launch__grid_size']] in ('320', '240', '160', '80'):
        def val(name):
            v = float(r[col[name]].replace(',', ''))
            u = units[col[name]]
            return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        sel.append({'grid': int(r[col['launch__grid_size']]), 'time_us': float(r[col['gpu__time_duration.sum']].replace(',', '')),
                    'dram_read': val('dram__bytes_read.sum'), 'dram_write': val('dram__bytes_write.sum'),
                    'tensor_pct': float(r[col['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']])})
if not sel:
    sys.exit('no rnn4 LSTM launch in the report')
full = [x for x in sel if x['grid'] == max(y['grid'] for y in sel)]
tr = sum(x['dram_read'] + x['dram_write'] for x in full) / len(full)
json.dump({'kernel': 'rc_tc_kernel<128,3,LSTM> rnn4 layer', 'launches_sampled': len(full), 'grid_ctas': full[0]['grid'],
           'dram_bytes_per_launch': tr, 'time_us_cold': sum(x['time_us'] for x in full) / len(full),
           'tensor_pipe_active_pct': sum(x['tensor_pct'] for x in full) / len(full), 'all': sel}, open(out, 'w'), indent=1)
print(open(out).read()[:600])
