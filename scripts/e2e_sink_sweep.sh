# e2e leg of bench.py under different sinks: output directory x write method x writer threads (one GPU)
for cfg in "/tmp pwrite 4" "/dev/shm pwrite 4" "/dev/shm mmap 8" "/tmp mmap 8" "/dev/shm pwrite 8" "/dev/shm mmap 4"; do
  set -- $cfg
  HB_BENCH_E2E_DIR=$1 HEYBUDDY_B200_SINK=$2 HB_BENCH_WRITERS=$3 python bench.py --no-cpu-baseline > gpurun_out/sweep.json 2>/dev/null
  python - "$cfg" <<PY
import json,sys
d=json.loads(open("gpurun_out/sweep.json").read().strip().splitlines()[-1]); e=d["e2e"]
print(sys.argv[1], "| e2e", round(e["value"]), "ms/step", round(e["ms_per_step"],2), "| sink GB/s", [round(x,2) for x in e["sink_gbs_per_rank"]], "| host mem", round(e["to_host_memory"]["value"]), "| value", round(d["value"]), flush=True)
PY
done
