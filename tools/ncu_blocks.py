"""Where the issue slots of a kernel go: consecutive SASS lines of an .ncu-rep grouped into runs of equal execution count,
with each run's share of the warp-instructions and of the stall samples (ncu --set full --import-source on capture)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []; kern = 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        kern += 1
        if kern > 1: break
        print(r[1]); continue
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
si = hdr.index('Warp Stall Sampling (All Samples)'); ie = hdr.index('Instructions Executed')
ex = [int(r[ie] or 0) for r in data]; sm = [int(r[si] or 0) for r in data]
tot, tots = sum(ex), sum(sm)
warps = max(ex)
print(f'warp-instructions {tot}  warps {warps}  per warp {tot / warps:.1f}  samples {tots}')
runs = []; s = 0
for i in range(1, len(ex) + 1):
    if i == len(ex) or abs(ex[i] - ex[s]) > 0.02 * max(ex[s], 1):
        runs.append((s, i - 1)); s = i
for a, b in runs:
    n = sum(ex[a:b + 1]); share = 100.0 * n / tot
    if share >= minshare:
        ops = {}
        for r in data[a:b + 1]:
            op = r[1].strip().split()[0]
            if op.startswith('@'): op = r[1].strip().split()[1]
            op = op.split('.')[0]; ops[op] = ops.get(op, 0) + 1
        top = ' '.join(f'{k}:{v}' for k, v in sorted(ops.items(), key=lambda t: -t[1])[:6])
        print(f'#{a:4d}-{b:4d} lines {b - a + 1:3d} exec/warp {ex[a] / warps:5.2f}  instr share {share:5.1f}%  stall share {100.0 * sum(sm[a:b + 1]) / tots:5.1f}%  {top}')
