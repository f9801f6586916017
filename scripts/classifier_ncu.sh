# stall reasons and pipe utilisation of the fused classifier step's kernels (one step after warm-up)
ncu --set full --import-source on --clock-control none -k regex:"fwd_tail|bwd_tail|gemm_tf32x3|finish_kernel|pre_kernel|head_train" --launch-skip 48 -c 8 -o gpurun_out/cls_step -f python scripts/classifier_time.py > gpurun_out/cls_ncu.log 2>&1
ncu -i gpurun_out/cls_step.ncu-rep --page raw --csv > gpurun_out/cls_step_raw.csv
