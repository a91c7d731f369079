import csv, sys, re
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
pats=sys.argv[2:] or ['gpu__time_duration.sum','dram__bytes_read.sum$','dram__bytes_write.sum$','registers_per_thread','occupancy','warps_active.avg.pct','issue_active.avg.pct','inst_executed.sum$','pipe_fp64','bank_conflict','stalled.*ratio','throughput.avg.pct','eligible','thread_inst_executed_per_inst','shared_mem','waves','l1tex__t_sector_hit','lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print('=====', r[hdr.index('Kernel Name')][:50], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for i,h in enumerate(hdr):
        if any(re.search(p,h) for p in pats):
            print(f"  {h:95s} {r[i]:>18s} {units[i]}")
