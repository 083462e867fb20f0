"""
Static SASS instruction counts of one kernel, bucketed by the source function (line ranges) they come from.

    python tools/sass_lines.py <cubin> <mangled kernel name> <file> lo:hi:name [lo:hi:name ...]

Uses `nvdisasm -c -gi` (line info incl. inlining). Counts are static (not weighted by execution).
"""
import collections
import re
import subprocess
import sys


def main():
    cubin, kern, fname = sys.argv[1:4]
    ranges = [(int(a), int(b), n) for a, b, n in (x.split(':') for x in sys.argv[4:])]
    txt = subprocess.run(['nvdisasm', '-c', '-gi', cubin], stdout=subprocess.PIPE, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(txt) if l.startswith('.text.' + kern + ':'))
    end = next((i for i in range(start + 1, len(txt)) if txt[i].startswith('.text.')), len(txt))
    cur = None
    per = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    for l in txt[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            f, ln = m.group(1).split('/')[-1], int(m.group(2))
            if f != fname and m.group(3) and m.group(3).split('/')[-1] == fname:
                ln, f = int(m.group(4)), fname
            cur = (f, ln)
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', l)
        if m and cur:
            b = cur[0]
            if cur[0] == fname:
                b = next((n for lo, hi, n in ranges if lo <= cur[1] <= hi), 'other')
            per[b] += 1
            ops[b][m.group(1)] += 1
    for b, n in per.most_common():
        print(f'{b:24s} {n:6d}  ' + ' '.join(f'{o}:{c}' for o, c in ops[b].most_common(12)))


if __name__ == '__main__':
    main()
