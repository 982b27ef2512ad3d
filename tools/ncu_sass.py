"""SASS-level view of an .ncu-rep: top instructions by stall samples, and every global memory instruction with its sector counts."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kern = 0
hdr = None; data = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        kern += 1
        if kern > 1: break
        print(r[1]); continue
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
si = hdr.index('Warp Stall Sampling (All Samples)'); ie = hdr.index('Instructions Executed')
sec = hdr.index('L2 Theoretical Sectors Global'); ideal = hdr.index('L2 Theoretical Sectors Global Ideal'); tag = hdr.index('L1 Tag Requests Global')
tot = sum(int(r[si] or 0) for r in data); toti = sum(int(r[ie] or 0) for r in data)
print('samples', tot, 'warp-instructions', toti, 'sass lines', len(data))
print('--- top by samples')
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][si] or 0))[:topn]:
    print(f'{int(r[si])*100/tot:5.1f}%  #{idx:4d} exec {int(r[ie] or 0):9d}  {r[1].strip()[:90]}')
print('--- global memory instructions (sectors, ideal, tag requests)')
for idx, r in enumerate(data):
    op = r[1].strip()
    if any(op.startswith(p) or (' ' + p) in op[:14] for p in ('LDG', 'STG', 'RED', 'ATOM')) or 'LDG' in op or 'STG' in op or 'REDG' in op or 'ATOMG' in op:
        print(f'#{idx:4d} exec {int(r[ie] or 0):9d} smp {int(r[si])*100/tot:5.1f}% sectors {r[sec]:>10} ideal {r[ideal]:>10} tags {r[tag]:>9}  {op[:80]}')
