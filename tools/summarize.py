"""Summaries of gpurun_out: bench lines and the per-hour-of-day kernel durations from the ncu launch list."""
import csv, json, sys, os
d = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out'
for f in ('bench_10m.json', 'bench_1m.json'):
    p = os.path.join(d, f)
    if os.path.exists(p) and os.path.getsize(p):
        for line in open(p):
            if line.startswith('{'):
                j = json.loads(line)
                print(f, 'value %.3e' % j['value'], 'e2e %.3e' % j['e2e']['value'], 'ms/day %.3f' % j['ms_per_step'],
                      {k: round(v, 4) for k, v in j['roofline']['per_kernel_ms'].items()}, 'frac %.3f' % j['roofline']['frac'], 'whole %.3f' % j['roofline']['whole_run_frac'])
p = os.path.join(d, 'launches.csv')
if os.path.exists(p):
    rows = list(csv.reader(open(p, errors='ignore')))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i
            break
    kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    seq = []
    for r in rows[start + 2:]:
        if len(r) <= mv: continue
        try: v = float(r[mv].replace(',', ''))
        except ValueError: continue
        seq.append((r[kn].split('(')[0].replace('void ', ''), v / 1e3))
    n, day = 0, []
    for name, v in seq:
        if name == 'k_sleep':
            n += 1
            if n == 3: day = []
        if n == 3: day.append((name, v))
        if n == 4: break
    hrs = [7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 0]
    k, tot = 0, 0.0
    out = []
    for name, v in day:
        tot += v
        if name.startswith('k_hour'): cur = 'h=%2d hour %.0f' % (hrs[k], v)
        elif name == 'k_resolve': cur += ' resolve %.0f' % v
        elif name == 'k_commit': out.append(cur + ' commit %.0f' % v); k += 1
        else: out.append('%s %.0f' % (name, v))
    print('one day under ncu (us, cold/serialised): total %.0f' % tot)
    print(' | '.join(out))
