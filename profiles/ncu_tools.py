"""Helpers to summarise ncu output (run here, no GPU needed).

    python profiles/ncu_tools.py launches gpurun_out/x_launches.csv
    python profiles/ncu_tools.py raw gpurun_out/x_prof.ncu-rep [kernel-substring]
    python profiles/ncu_tools.py stalls gpurun_out/x_prof.ncu-rep kernel-substring
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__average_warp_latency_issue_stalled_barrier.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        ]


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
        k = row['Kernel Name'].split('(')[0][:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print('%-72s %6s %12s %10s %6s' % ('kernel', 'n', 'total us', 'avg us', 'share'))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-72s %6d %12.1f %10.1f %5.1f%%' % (k, n, t, t / n, 100 * t / tot))
    print('total %.1f us' % tot)


def raw(path, sub=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        if sub and sub not in name:
            continue
        print('---', name[:100])
        for k in KEYS:
            if k in hdr:
                print('  %-85s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))


def stalls(path, sub, top=40):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '-k', 'regex:' + sub],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    data = []
    for r in rows:
        if 'Source' in r and '# Samples' in ' '.join(r) or (hdr is None and 'Source' in r):
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    if not hdr:
        print(out[:2000])
        return
    si = [i for i, h in enumerate(hdr) if h.startswith('Warp Stall Sampling (All')]
    si = si[0] if si else [i for i, h in enumerate(hdr) if 'Sampl' in h][0]
    src = hdr.index('Source')
    tot = sum(float(r[si] or 0) for r in data)
    print('total samples', tot, '| column:', hdr[si])
    for r in sorted(data, key=lambda r: -float(r[si] or 0))[:top]:
        print('%7.0f %5.1f%%  %s' % (float(r[si] or 0), 100 * float(r[si] or 0) / max(tot, 1),
                                     r[src][:150]))


if __name__ == '__main__':
    cmd = sys.argv[1]
    if cmd == 'launches':
        launches(sys.argv[2])
    elif cmd == 'raw':
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    elif cmd == 'phases':
        pass   # defined below
    else:
        stalls(sys.argv[2], sys.argv[3])


def phases(path, sub):
    """Stall samples / executed instructions between barrier-type SASS markers."""
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '-k', 'regex:' + sub],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    his = [i for i, r in enumerate(rows) if 'Source' in r]
    for hn, hi in enumerate(his):
        hdr = rows[hi]
        end = his[hn + 1] if hn + 1 < len(his) else len(rows)
        si = [i for i, h in enumerate(hdr) if h.startswith('Warp Stall Sampling (All')][0]
        ex, src = hdr.index('Instructions Executed'), hdr.index('Source')
        data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
        tot = sum(float(r[si] or 0) for r in data)
        print('=== launch %d: %d SASS lines, %d samples' % (hn, len(data), tot))
        acc = inst = 0.0
        start = 0
        for i, r in enumerate(data):
            acc += float(r[si] or 0)
            inst += float(r[ex] or 0)
            s = r[src]
            if 'BAR.SYNC' in s or 'UTCBAR' in s or 'SYNCS.PHASECHK' in s or i == len(data) - 1:
                if acc > tot * 0.004:
                    print('%5d-%5d %7.0f %5.1f%%  inst=%10.0f  .. %s' % (
                        start, i, acc, 100 * acc / max(tot, 1), inst, s.strip()[:46]))
                acc = inst = 0.0
                start = i + 1


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'phases':
    phases(sys.argv[2], sys.argv[3])
