"""One simulated day from an ncu launch list (gpu__time_duration + dram bytes): per hour-of-day k_hour / k_commit time (us) and DRAM MB."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors='ignore')) if len(r) > 10]
h = rows[0]; kn = h.index('Kernel Name'); mn = h.index('Metric Name'); mv = h.index('Metric Value'); idc = h.index('ID')
L = collections.OrderedDict()
for r in rows[1:]:
    L.setdefault(int(r[idc]), {'k': r[kn].split('(')[0].replace('void ', '')})[r[mn]] = float(r[mv].replace(',', ''))
seq = list(L.values())
# find the 3rd k_sleep -> one whole day follows
n = 0; day = []
for e in seq:
    if e['k'] == 'k_sleep':
        n += 1
        if n == 3: day = []
    if n == 3: day.append(e)
    if n == 4: break
hrs = [7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 0]
k = 0; tot = 0; out = []
for e in day:
    t = e.get('gpu__time_duration.sum', 0) / 1e3; mb = (e.get('dram__bytes_read.sum', 0) + e.get('dram__bytes_write.sum', 0)) / 1e6
    tot += t
    if e['k'].startswith('k_hour'): cur = 'h=%2d hour %3.0fus %4.0fMB' % (hrs[k], t, mb)
    elif e['k'].startswith('k_commit'): out.append(cur + ' | commit %3.0fus %4.0fMB' % (t, mb)); k += 1
    else: out.append('%s %.0fus %.0fMB' % (e['k'], t, mb))
print('day total %.0f us' % tot); print('\n'.join(out))
# --json FILE: DRAM bytes of the average active-hour pass of that day (all 18 k_hour + 18 k_commit launches), for bench.py's roofline.traffic
if '--json' in sys.argv:
    import json
    kh = [e for e in day if e['k'].startswith('k_hour')]; kc = [e for e in day if e['k'].startswith('k_commit')]
    byt = lambda e: e.get('dram__bytes_read.sum', 0) + e.get('dram__bytes_write.sum', 0)
    unit = 1.0
    j = {'k_hour': {'dram_bytes_per_launch': sum(map(byt, kh)) * unit / len(kh), 'launches_captured': len(kh)},
         'k_commit': {'dram_bytes_per_launch': sum(map(byt, kc)) * unit / len(kc), 'launches_captured': len(kc)},
         'workload': '10m',
         'source': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on bench.py --steps 2 --warmup 3: every k_hour / k_commit launch of one simulated day (day 3), B200; profiles/r02d_launches_10m.csv'}
    j['pass_dram_bytes_per_launch'] = j['k_hour']['dram_bytes_per_launch'] + j['k_commit']['dram_bytes_per_launch']
    # the 16 movement hours (h = 7..22) alone: the first 16 of the day's 18 k_hour / k_commit pairs (h = 23 and h = 0 follow)
    j['movement_pass_dram_bytes_per_launch'] = (sum(map(byt, kh[:16])) + sum(map(byt, kc[:16]))) / 16.0
    j['movement_pass_us'] = (sum(e.get('gpu__time_duration.sum', 0) for e in kh[:16] + kc[:16])) / 16e3
    json.dump(j, open(sys.argv[sys.argv.index('--json') + 1], 'w'), indent=1)

