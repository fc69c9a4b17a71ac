"""Turn the ncu artefacts a gpurun call brought back into the text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r01_s3.csv profiles/r01_launch_shares_s3.txt
    python tools/ncu_summary.py full gpurun_out/prof_r01_s3.ncu-rep profiles/r01_ncu_full_summary_s3.txt \
        profiles/r01_traffic.json
    python tools/ncu_summary.py opcodes gpurun_out/prof_r01_s3.ncu-rep k_throw profiles/r01_k_throw_opcodes_s3.txt

(ncu must be on PATH; the .ncu-rep files are read here, not on the GPU box.)
"""
import collections
import csv
import io
import json
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum',
    'lts__t_sector_hit_rate.pct',
]


def _rows(cmd):
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    return [r for r in csv.reader(io.StringIO(out)) if r]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith('==')]
    head = rows[0]
    iname, ival, imet = head.index('Kernel Name'), head.index('Metric Value'), head.index('Metric Name')
    iunit = head.index('Metric Unit')
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if r[imet] != 'gpu__time_duration.sum':
            continue
        v = float(r[ival].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[iunit], 1e-6)
        tot[r[iname]] += v
        cnt[r[iname]] += 1
    s = sum(tot.values())
    with open(dst, 'w') as f:
        f.write('ncu launch list (%s; cold-cache, serialised): share of device time\n' % src)
        for k, v in tot.most_common():
            f.write('%-78s n=%3d total=%9.3f ms  %5.1f%%\n' % (k[:78], cnt[k], v, 100 * v / s))
        ex = {k: v for k, v in tot.items() if any(t in k for t in ('k_throw', 'k_reads', 'k_counts'))}
        se = sum(ex.values())
        f.write('\nshares among the exposure kernels:\n')
        for k, v in sorted(ex.items(), key=lambda kv: -kv[1]):
            f.write('  %-60s %5.1f%%\n' % (k[:60], 100 * v / se))


def full(rep, dst, traffic_json=None):
    rows = _rows(['ncu', '-i', rep, '--page', 'raw', '--csv'])
    head, units = rows[0], rows[1]
    traffic = {}
    with open(dst, 'w') as f:
        f.write('ncu --set full --clock-control none (%s)\n' % rep)
        for v in rows[2:]:
            name = v[head.index('Kernel Name')]
            f.write('----\nKernel Name  %s\n' % name[:110])
            for m in METRICS:
                if m in head:
                    f.write('%-76s %s %s\n' % (m, v[head.index(m)], units[head.index(m)]))
            for i, n in enumerate(head):
                if 'smsp__average_warps_issue_stalled' in n and n.endswith('per_issue_active.ratio'):
                    try:
                        if float(v[i]) >= 0.3:
                            f.write('  stall %-60s %s\n' % (n.split('issue_stalled_')[1].split('_per_issue')[0], v[i]))
                    except ValueError:
                        pass
            rd = float(v[head.index('dram__bytes_read.sum')]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[
                units[head.index('dram__bytes_read.sum')]]
            wr = float(v[head.index('dram__bytes_write.sum')]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[
                units[head.index('dram__bytes_write.sum')]]
            key = 'k_throw' if 'k_throw' in name else ('k_reads' if 'k_reads' in name else (
                'k_counts' if 'k_counts' in name else name[:20]))
            traffic[key] = int(rd + wr)
    if traffic_json:
        traffic['source'] = ('ncu --set full --clock-control none, %s (dram__bytes_read.sum + '
                             'dram__bytes_write.sum per launch, workload c4)' % dst)
        with open(traffic_json, 'w') as f:
            json.dump(traffic, f, indent=1)


def opcodes(rep, kernel, dst):
    rows = _rows(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel])
    head = next(r for r in rows if 'Source' in r and 'Instructions Executed' in r)
    ia, isrc, isamp = head.index('Instructions Executed'), head.index('Source'), head.index('# Samples')
    body = [r for r in rows if len(r) > ia and r[ia].isdigit()]
    seen, uniq = set(), []
    for r in body:                      # the page lists a kernel once per captured launch
        if r[0] in seen:
            break
        seen.add(r[0])
        uniq.append(r)
    h, hs = collections.Counter(), collections.Counter()
    for r in uniq:
        op = r[isrc].split()
        o = op[1] if op[0].startswith('@') else op[0]
        o = o if o.startswith('MUFU') else o.split('.')[0]
        if o == 'IMAD' and '.WIDE' in r[isrc]:
            o = 'IMAD.WIDE'
        h[o] += int(r[ia])
        hs[o] += int(r[isamp])
    tot = sum(h.values())
    with open(dst, 'w') as f:
        f.write('executed SASS opcodes of %s (%s, source page)\ntotal warp-instructions %d\n' % (kernel, rep, tot))
        for o, n in h.most_common(45):
            f.write('%-12s %12d %5.1f%%  samples %d\n' % (o, n, 100 * n / tot, hs[o]))


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'opcodes': opcodes}[sys.argv[1]](*sys.argv[2:])
