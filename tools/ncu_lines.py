"""Per-source-line executed warp instructions and stall samples of the first kernel in an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; fname = None; kern = 0; data = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        kern += 1
        if kern > 1: break
        continue
    if r and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == '-': data.append((fname, r))
si = hdr.index('Warp Stall Sampling (All Samples)'); ie = hdr.index('Instructions Executed')
tot = sum(int(r[si] or 0) for _, r in data); toti = sum(int(r[ie] or 0) for _, r in data)
print('samples', tot, 'warp-instr', toti)
for f, r in sorted(data, key=lambda t: -int(t[1][ie] or 0))[:topn]:
    print(f'{int(r[ie] or 0)*100/toti:5.1f}% ins {int(r[si])*100/tot:5.1f}% smp  {f}:{r[0]}: {r[1].strip()[:110]}')
