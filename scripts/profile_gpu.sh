#!/bin/bash
# Run under gpurun (one GPU).  Produces, in gpurun_out/:
#   launches.csv          every launch of our kernels with its device time (ncu, cold-cache, serialised)
#   <name>.ncu-rep        one --set full capture per heavy kernel
CFG=${1:-C}
shift
KERNELS=${@:-"spmm_f32 resid_kernel xb_tc gram_tc onehot_step"}
K='regex:spmm_f32|onehot_step|row_kurtosis|batch_kurtosis|resid_kernel|gram_|xb_|split_f16|perm_stats|absmax|obs_hist|cell_fdr|colsum|scale_kernel|bfs_|permute_'
BENCH="python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
for name in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$name -s 1 -c 1 -f -o gpurun_out/$name $BENCH > gpurun_out/ncu_$name.log 2>&1
done
ls -la gpurun_out/
