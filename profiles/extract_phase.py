"""Read the `ncu --set full` report of the persistent grouped kernel (via `ncu -i X.ncu-rep --page raw --csv`) and write its per-launch
summary (duration, tensor pipe, DRAM / L2 traffic) to profiles/r01_phase_traffic.json plus the raw page to profiles/r01_ncu_phase_raw.csv."""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else 'profiles/r01_phase_traffic.json'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
open('profiles/r01_ncu_phase_raw.csv', 'w').write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
sel = []
for r in rows[2:]:
    if 'rc_tc_phase' not in r[col['Kernel Name']]:
        continue

    def val(name):
        return float(r[col[name]].replace(',', '')) * scale.get(units[col[name]], 1)
    sel.append({'kernel': r[col['Kernel Name']].split('(')[0], 'grid': int(r[col['launch__grid_size']]),
                'time_us_cold': val('gpu__time_duration.sum'),
                'dram_read': val('dram__bytes_read.sum'), 'dram_write': val('dram__bytes_write.sum'),
                'tma_l2_to_sm_bytes': val('l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum'),
                'tensor_pipe_active_pct': val('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
                'l2_hit_pct': val('lts__t_sector_hit_rate.pct'),
                'lts_throughput_pct': val('lts__throughput.avg.pct_of_peak_sustained_elapsed'),
                'registers': val('launch__registers_per_thread')})
if not sel:
    sys.exit('no rc_tc_phase launch in the report')
# one frame = the launches of the report (phase 1, phase 2, vision updater); bench.py's roofline.traffic is per launch
tr = sum(x['dram_read'] + x['dram_write'] for x in sel) / len(sel)
json.dump({'kernel': sel[0]['kernel'], 'launches_sampled': len(sel), 'dram_bytes_per_launch': tr,
           'note': 'launches of one frame: rnn4+rnn2 | rnn6+rnn3+rnn7+rnn8 | vision updater (rnn4+rnn6 on the low-confidence rows); '
                   'algorithmic weight bytes of the three together = 243 MB (split fp16 hi+lo = 4 B per weight)',
           'all': sel}, open(out, 'w'), indent=1)
print(open(out).read()[:1500])
