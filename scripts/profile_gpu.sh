#!/bin/bash
# Run under gpurun (one GPU).  Produces, in gpurun_out/:
#   launches.csv          every launch of our kernels (namespaces cna:: and tc::) with its device time
#                         (ncu, cold-cache, serialised: compare shares, not absolutes)
#   <name>.ncu-rep        one --set full capture per heavy kernel
#   ncu_<name>.txt        scripts/ncu_summary.py over that capture
CFG=${1:-C}
shift
KERNELS=${@:-"spmm_f32_kernel resid_lin_kernel xb_tc_kernel gram_tc_kernel sym_eig_top_kernel onehot_step_kernel"}
K='regex:spmm_|onehot_step|resid_|gram_|xb_tc|sym_eig|perm_stats|minp_|median|select_|absmax|obs_hist|cell_fdr|fdr_|mt_stream|attempts_|chunk_scan|final_state|rank_kernel|colsum|scale_kernel|bfs_|permute_csr|row_kurtosis|batch_kurtosis|split_f16|kurt'
BENCH="python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 2000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
for name in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$name -s 1 -c 1 -f -o gpurun_out/$name $BENCH > gpurun_out/ncu_$name.log 2>&1
  python scripts/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/ncu_$name.txt 2>&1
done
ls -la gpurun_out/
