# per-kernel times and pipe utilisation of the pitch-shift kernels (one group of 512 clips)
ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ps_ -c 4 --csv --log-file gpurun_out/k9_pitch_ncu.csv python scripts/k9_time.py > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/k9_pitch_ncu.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
H=rows[h]; ix={n:i for i,n in enumerate(H)}
cur=None
for r in rows[h+1:]:
    if len(r)<len(H): continue
    key=(r[ix["ID"]], r[ix["Kernel Name"]][:34])
    if key!=cur: print(); print(key[1], end=": "); cur=key
    print(r[ix["Metric Name"]].split(".")[0].replace("__","_")[-28:], r[ix["Metric Value"]], r[ix["Metric Unit"]], end=" | ")
print()
PY
