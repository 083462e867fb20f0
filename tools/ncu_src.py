"""Per-source-line executed instructions / stall samples of an ncu capture (needs -lineinfo and --import-source on).

    python tools/ncu_src.py <rep> <file name> [lo hi]      lines lo..hi of that file (executed warp instructions, samples)
"""
import csv, io, subprocess, sys


def rows_of(rep):
    out = subprocess.run(['ncu', '-i', rep, '--csv', '--page', 'source', '--print-source', 'cuda,sass'],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    cur, hdr = None, None
    for r in csv.reader(io.StringIO(out)):
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
                yield cur, int(r[0]), r[1].rstrip(), int(d['Instructions Executed']), int(d['# Samples'])
            except (ValueError, KeyError):
                pass


def main():
    rep, fname = sys.argv[1], sys.argv[2]
    lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 10 ** 9)
    data = list(rows_of(rep))
    tot = sum(x[3] for x in data) or 1
    tots = sum(x[4] for x in data) or 1
    for f, ln, src, inst, samp in data:
        if f == fname and lo <= ln <= hi and (inst or samp):
            print(f'{ln:5d} inst {100. * inst / tot:6.2f}% samp {100. * samp / tots:6.2f}% | {src[:110]}')


if __name__ == '__main__':
    main()
