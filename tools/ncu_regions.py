"""Summarise an ncu report of one kernel: headline metrics + instruction share / active lanes per SASS block."""
import csv, subprocess, sys, io
rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, it, isrc, iss = (hdr.index(k) for k in ('Instructions Executed', 'Thread Instructions Executed', 'Source', '# Samples'))
data = []
for r in rows[2:]:
    if len(r) > it and r[ia].isdigit():
        data.append(r)
    elif data:
        break
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[iss]) for r in data)
print("total warp instructions", tot)
for b in range(0, len(data), B):
    blk = data[b:b + B]
    a = sum(int(r[ia]) for r in blk); t = sum(int(r[it]) for r in blk); s = sum(int(r[iss]) for r in blk)
    if a < tot * 0.002:
        continue
    ops = ' '.join(sorted(set(r[isrc].split()[1 if r[isrc].startswith('@') else 0].split('.')[0] for r in blk if r[isrc]))[:12])
    print(f"{b:5d} inst={a / tot * 100:5.2f}% thr={t / max(a, 1):5.1f} smp={s / max(tots,1) * 100:5.2f}%  {ops[:100]}")
