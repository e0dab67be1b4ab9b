import csv, sys, json
src, dst = sys.argv[1], sys.argv[2]
rows=list(csv.reader(open(src)))
h=rows[0]
keys=['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active']
idx={k:h.index(k) for k in keys}
out=[]
for r in rows[2:]:
    out.append({k:(r[idx[k]][:100] if k=='Kernel Name' else float(r[idx[k]])) for k in keys})
units={k:rows[1][idx[k]] for k in keys if k!='Kernel Name'}
json.dump({'source': 'ncu --set full --clock-control none (cold-cache, serialised replays: compare shares, not absolutes)', 'units':units,'kernels':out}, open(dst,'w'), indent=1)
print(len(out), 'kernels ->', dst)
