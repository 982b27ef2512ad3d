"""Read an .ncu-rep here (no GPU): key metrics, stall-reason totals and the hottest source lines."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_registers',
        'sm__maximum_warps_per_active_cycle_pct', 'lts__t_bytes.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum']
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            print(f'{k} = {r[hdr.index(k)]}')
    for k, v in zip(hdr, r):
        if 'smsp__average_warps_issue_stalled' in k and '_not_issued' not in k and k.endswith('.ratio'):
            try:
                if float(v) > 0.3: print(f'   {k.replace("smsp__average_warps_issue_stalled_","stall ").replace("_per_issue_active.ratio","")} = {v}')
            except ValueError: pass
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
data = []
for r in rows:
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
if hdr:
    si = hdr.index('Warp Stall Sampling (All Samples)'); ie = hdr.index('Instructions Executed')
    tot = sum(int(r[si] or 0) for r in data); toti = sum(int(r[ie] or 0) for r in data)
    print('total samples', tot, 'warp instr', toti)
    for r in sorted(data, key=lambda r: -int(r[si] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
        print(f'{int(r[si])*100/tot:5.1f}% smp {int(r[ie] or 0)*100/toti:5.1f}% ins  L{r[0]}: {r[1].strip()[:130]}')
