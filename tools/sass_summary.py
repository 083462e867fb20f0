"""SASS evidence for profiles/: per kernel of the built library, the instruction count, the mnemonic mix that matters
for this path (FP64 pipe, shared memory, barriers, bulk async copies, shuffles, reductions) and the lines that show the
TMA / bulk-copy and mbarrier instructions.

    python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'spitfire_b200/libgriffon_b200.so'
txt = subprocess.run(['cuobjdump', '-sass', lib], stdout=subprocess.PIPE, text=True).stdout.splitlines()
KEEP = ['k_jacILi8', 'k_jac4ILi512', 'k_rates', 'k_btddod_invertILi2ELi7', 'k_btddod_solve_inv', 'k_btddod_factorizeILi16ELi4',
        'k_btddod_solveILi2', 'k_block_max_real_eig', 'k_flamelet_prepass', 'k_newton_tail', 'k_esdirk_finish',
        'k_stage_begin']
GROUPS = [('FP64', r'^(DFMA|DMUL|DADD|DSETP|DMNMX)'), ('MUFU', r'^MUFU'), ('LDS', r'^LDS'), ('STS', r'^STS'),
          ('LDG', r'^(LDG|LD\.)'), ('STG', r'^(STG|ST\.)'), ('BAR', r'^BAR'), ('SHFL', r'^SHFL'), ('REDUX', r'^REDUX'),
          ('bulk copy (TMA)', r'^(UBLKCP|UBLKRED|UTMALDG|UTMASTG)'), ('LDGSTS', r'^LDGSTS'),
          ('mbarrier', r'^(SYNCS|ARRIVES)'), ('fence/membar', r'^(FENCE|MEMBAR)'), ('HMMA/DMMA/tcgen05', r'^(HMMA|DMMA|UTCHMMA|UTCMMA)')]
cur, per, lines = None, collections.OrderedDict(), collections.defaultdict(list)
for l in txt:
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = m.group(1) if any(k in m.group(1) for k in KEEP) else None
        if cur:
            per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);', l)
    if not m:
        continue
    op = m.group(2)
    per[cur]['total'] += 1
    for name, pat in GROUPS:
        if re.match(pat, op):
            per[cur][name] += 1
            if name in ('bulk copy (TMA)', 'mbarrier') and len(lines[(cur, name)]) < 4:
                lines[(cur, name)].append(f'/*{m.group(1)}*/ {op}{m.group(3)}')
print(f'# cuobjdump -sass {lib}: static instruction counts per kernel (tools/sass_summary.py)')
for k, c in per.items():
    print(f'\n{k}\n  {c["total"]} instructions: ' + ', '.join(f'{n} {c[n]}' for n, _ in GROUPS if c[n]))
    for name in ('bulk copy (TMA)', 'mbarrier'):
        for ln in lines[(k, name)]:
            print(f'    {ln}')
