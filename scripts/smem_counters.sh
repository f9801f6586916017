#!/bin/bash
# shared-memory pipe counters of the grouped embed kernels (LSU wavefronts / bank conflicts by op, tensor-core operand wavefronts)
M=l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:"tcg_block|tail_layer" -c 8 --csv --log-file gpurun_out/smem.csv python scripts/embed_time.py ${1:-8192} > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/smem.csv')) if len(r)>8]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[iid],r[ik][-58:-20]),{})[r[im].replace('l1tex__data_','').replace('_mem_shared','').replace('pipe_lsu_','')]=r[iv]
for k,v in d.items():
    print(k[1], {a[:40]:b for a,b in v.items()})
PY
