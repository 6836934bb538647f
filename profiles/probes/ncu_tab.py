import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE).stdout.decode()
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
def col(name): return hdr.index(name) if name in hdr else None
want = [('Kernel Name','k'),('Grid Size','grid'),('gpu__time_duration.sum','us'),('dram__bytes_read.sum','dR'),('dram__bytes_write.sum','dW'),
 ('lts__t_bytes.sum','l2B'),('lts__t_sector_hit_rate.pct','l2hit'),('sm__warps_active.avg.pct_of_peak_sustained_active','occ%'),
 ('launch__registers_per_thread','regs'),('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','stall_ls'),
 ('l1tex__t_bytes.sum','l1B'),('dram__throughput.avg.pct_of_peak_sustained_elapsed','dram%'),('lts__throughput.avg.pct_of_peak_sustained_elapsed','lts%'),
 ('sm__throughput.avg.pct_of_peak_sustained_elapsed','sm%'),('l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1%')]
idx = [(lab, col(n)) for n, lab in want]
print(' | '.join(l for l, i in idx if i is not None))
for r in rows[2:]:
    vals = []
    for lab, i in idx:
        if i is None: continue
        v = r[i]
        if lab == 'k': v = v.split('(')[0].replace('void agx::','')[:28]
        vals.append(f'{v} {units[i]}' if lab in ('us','dR','dW','l2B','l1B') else v)
    print(' | '.join(vals))
