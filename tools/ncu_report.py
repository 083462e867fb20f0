"""Summarise an ncu capture: headline metrics, stall mix, instruction/sample shares per kernel phase (phases are found
from the `// ----` marker comments of the CUDA source) and the hottest source lines.

    python tools/ncu_report.py gpurun_out/prof.ncu-rep spitfire_b200/csrc/gb_jac.cu > profiles/summary.txt
"""
import collections
import csv
import io
import subprocess
import sys


def ncu_csv(rep, *args):
    out = subprocess.run(['ncu', '-i', rep, '--csv'] + list(args), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, srcfile = sys.argv[1], sys.argv[2]
    raw = ncu_csv(rep, '--page', 'raw')
    h, u, v = raw[0], raw[1], raw[2]
    keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
            'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
            'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second']
    print('== headline metrics ==')
    for k in keys:
        if k in h:
            i = h.index(k)
            print(f'{k:85s} {v[i]:>18s} {u[i]}')
    rows = ncu_csv(rep, '--page', 'source', '--print-source', 'cuda,sass')
    cur, hdr, data = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr and r[0].isdigit():
            d = dict(zip(hdr, r))
            try:
                data.append(dict(file=cur, line=int(r[0]), src=r[1].strip()[:100], inst=int(d['Instructions Executed']),
                                 samp=int(d['# Samples']),
                                 stalls={k[6:]: int(d[k]) for k in d if k.startswith('stall_') and '(' not in k}))
            except (ValueError, KeyError):
                pass
    tot = sum(x['inst'] for x in data) or 1
    tots = sum(x['samp'] for x in data) or 1
    mix = collections.Counter()
    for x in data:
        mix.update(x['stalls'])
    print('\n== warp stall mix (% of samples) ==')
    print({k: round(100 * n / tots, 1) for k, n in mix.most_common() if n})
    src = open(srcfile).read().splitlines()
    marks = [(1, 'file head / device functions')]
    for i, line in enumerate(src):
        s = line.strip()
        if s.startswith('// ----') and len(s) > 12 and any(c.isalpha() for c in s):
            marks.append((i + 1, s.strip('/- ')[:60]))
        elif s.startswith('__global__') or (s.startswith('__device__') and '(' in s) or s.startswith('template <int G>'):
            marks.append((i + 1, 'fn: ' + s[:70]))
    marks.append((len(src) + 1, 'end'))
    agg = collections.OrderedDict()
    base = srcfile.split('/')[-1]
    for x in data:
        if x['file'] != base:
            key = x['file']
        else:
            key = '?'
            for (l0, n0), (l1, _) in zip(marks[:-1], marks[1:]):
                if l0 <= x['line'] < l1:
                    key = f'{l0:4d} {n0}'
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += x['inst']
        a[1] += x['samp']
        a[2] += x['stalls'].get('barrier', 0)
    print('\n== shares by phase (instructions executed / stall samples / of which barrier) ==')
    for k, a in agg.items():
        if a[0] or a[1]:
            print(f'{k:76s} inst {100 * a[0] / tot:5.1f}%  samples {100 * a[1] / tots:5.1f}%  barrier {100 * a[2] / tots:5.1f}%')
    print('\n== hottest source lines by samples ==')
    for x in sorted(data, key=lambda x: -x['samp'])[:25]:
        print(f"{x['file']:16s} {x['line']:5d} samples {100 * x['samp'] / tots:5.2f}% inst {100 * x['inst'] / tot:5.2f}% | {x['src']}")


if __name__ == '__main__':
    main()
