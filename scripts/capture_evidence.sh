#!/bin/bash
# Evidence capture on the GPU box (one gpurun call, one GPU): the plain run first, then the ncu launch list and the
# --set full capture of the heavy kernels of the SAME command (B200_PROFILING.md).  Outputs under gpurun_out/.
set -u
CMD="python bench.py --steps 2 --warmup 3 --cpu-pairs 0 --no-e2e --no-graph --no-others"
mkdir -p gpurun_out
$CMD > gpurun_out/ev_plain.json 2> gpurun_out/ev_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ev_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tau_kernel|round1|sparse_kernel|nms_rounds|select_kernel|warp_homography|sample_|prep_kernel|nn_top2|resolve_kernel|rescan_kernel|gate_kernel|pairs_kernel" -c 600 --csv --log-file gpurun_out/ev_launches.csv $CMD > gpurun_out/ev_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'round1_|sample_planes|nn_top2_kernel|prep_kernel|sparse_kernel' -s 40 -c 10 -o gpurun_out/ev_full $CMD > gpurun_out/ev_ncu_full.log 2>&1
ls -la gpurun_out/ev_*
