"""SASS summary of the in-tree shared library: occurrences per kernel of the Blackwell-native mnemonics (B200_PROFILING.md).
Usage: python profiles/sass_summary.py [out.txt]   (runs cuobjdump -sass on robustcap_b200/csrc/librobustcap_b200.so)"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(REPO, 'robustcap_b200', 'csrc', 'librobustcap_b200.so')
cols = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'UTMALDG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'LDGSTS', 'HMMA', 'FFMA', 'MUFU', 'STL', 'LDL', 'MEMBAR', 'RED.E']
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
counts, name = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[name] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and name:
        op = m.group(1)
        for c in cols:
            if c == 'UTCHMMA.2CTA':
                hit = op.startswith('UTCHMMA') and '2CTA' in op
            elif c == 'UTCHMMA':
                hit = op.startswith('UTCHMMA') and '2CTA' not in op
            else:
                hit = op.startswith(c)
            if hit:
                counts[name][c] += 1
out = ['# SASS summary of robustcap_b200/csrc/librobustcap_b200.so (cuobjdump -sass, sm_100a) — profiles/sass_summary.py',
       '# UTCHMMA = tcgen05.mma (.2CTA: cta_group::2), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk',
       '# (TMA 1-D bulk copy), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async, STL / LDL = register spills', '',
       '%-100s' % 'kernel' + ''.join('%13s' % c for c in cols)]
for k, v in counts.items():
    if any(v[c] for c in cols):
        out.append('%-100s' % k[:100] + ''.join('%13d' % v[c] for c in cols))
res = '\n'.join(out) + '\n'
if len(sys.argv) > 1:
    open(sys.argv[1], 'w').write(res)
else:
    print(res)
