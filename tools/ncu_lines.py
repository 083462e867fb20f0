"""Per-source-line instruction / stall-sample shares of an ncu capture, for a line range of one file.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep gb_jac.cu 142 370 [min_pct]
"""
import csv, io, subprocess, sys, collections


def load(rep):
    out = subprocess.run(['ncu', '-i', rep, '--csv', '--page', 'source', '--print-source', 'cuda,sass'],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
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
    return data


if __name__ == '__main__':
    rep, fname, l0, l1 = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    minpct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.2
    data = load(rep)
    tot = sum(x['inst'] for x in data) or 1
    tots = sum(x['samp'] for x in data) or 1
    files = collections.Counter()
    for x in data:
        files[x['file']] += x['inst']
    print({k: round(100 * v / tot, 2) for k, v in files.most_common()})
    for x in data:
        if x['file'] == fname and l0 <= x['line'] <= l1 and (100 * x['inst'] / tot >= minpct or 100 * x['samp'] / tots >= minpct):
            top = sorted(x['stalls'].items(), key=lambda kv: -kv[1])[:2]
            print(f"{x['line']:5d} inst {100*x['inst']/tot:5.2f}% samp {100*x['samp']/tots:5.2f}% {str(top):48s}| {x['src']}")
