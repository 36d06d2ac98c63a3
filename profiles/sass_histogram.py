"""SASS evidence per kernel of libgossipnet_b200.so (run here, no GPU needed):
    python profiles/sass_histogram.py > profiles/r2_sass_histogram.txt
Counts the Blackwell-specific mnemonics per kernel: UTC*MMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG (cp.async.bulk.tensor loads / stores, tensor-map TMA), UBLKCP (cp.async.bulk),
LDGSTS (cp.async), SYNCS (mbarrier), RED / ATOMG."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'gossipnet_b200', 'csrc', 'libgossipnet_b200.so')
KEYS = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'UTCBAR', 'HMMA',
        'FFMA', 'RED', 'ATOMG', 'STG', 'LDG', 'STS', 'LDS', 'MUFU']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    kern, hist, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r'\(.*', '', kern)
            hist[kern] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and kern:
            op = m.group(1)
            hist[kern]['_all'] += 1
            for k in KEYS:
                if op.startswith(k):
                    hist[kern][k] += 1
                    total[k] += 1
    print('%-52s %6s  %s' % ('kernel', 'instr', ' '.join('%7s' % k for k in KEYS)))
    for kern, c in hist.items():
        if not any(c[k] for k in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'LDGSTS')) and '--all' not in sys.argv:
            continue
        print('%-52s %6d  %s' % (kern.replace('gn::', '')[:52], c['_all'], ' '.join('%7d' % c[k] for k in KEYS)))
    print('%-52s %6s  %s' % ('whole library', '', ' '.join('%7d' % total[k] for k in KEYS)))


if __name__ == '__main__':
    main()
