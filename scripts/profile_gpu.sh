#!/bin/bash
# Run under gpurun (one GPU).  Produces, in gpurun_out/:
#   launches.csv   every launch of our kernels with its device time (ncu, cold-cache, serialised)
#   spmm.ncu-rep / nullhist.ncu-rep   one --set full capture of the two heaviest kernels
CFG=${1:-C}
K='regex:spmm_f32|onehot_step|row_kurtosis|batch_kurtosis|resid_kernel|gram_|xb_|null_hist|perm_stats|absmax|obs_hist|cell_fdr|colsum|scale_kernel'
BENCH="python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_f32 -s 2 -c 1 -f -o gpurun_out/spmm $BENCH > gpurun_out/ncu_spmm.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:xb_|null_hist' -s 1 -c 1 -f -o gpurun_out/nullhist $BENCH > gpurun_out/ncu_nullhist.log 2>&1
ls -la gpurun_out/
