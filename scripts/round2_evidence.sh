set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"colored_bases|augment_fast|tcg_block|tail_layer|gather_slots" --launch-skip 33 -c 11 -o gpurun_out/r2_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log | cut -c1-200
ncu -i gpurun_out/r2_step.ncu-rep --page raw --csv > gpurun_out/r2_raw.csv 2>/dev/null
bash scripts/smem_counters.sh 8192 > gpurun_out/r2_smem_counters.txt 2>&1
tail -9 gpurun_out/r2_smem_counters.txt | cut -c1-300
