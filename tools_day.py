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
    elif e['k'] == 'k_commit': out.append(cur + ' | commit %3.0fus %4.0fMB' % (t, mb)); k += 1
    else: out.append('%s %.0fus %.0fMB' % (e['k'], t, mb))
print('day total %.0f us' % tot); print('\n'.join(out))
