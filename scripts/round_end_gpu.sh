set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r1b_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b_smoke.log 2>&1; tail -2 gpurun_out/r1b_smoke.log
python bench.py > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1b_bench_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"augment_fast|mel_kernel|tcg_block|tc_block|gather_slots" --launch-skip 27 -c 9 -o gpurun_out/r1b_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log | cut -c1-200
cat gpurun_out/r1b_pytest_gpu.log; cut -c1-400 gpurun_out/r1b_bench_n1.json
