#!/bin/bash
# Round-2 ncu evidence (run on the GPU box: gpurun -- 'bash tools/run_profiles.sh'):
#   per workload: a plain run first (must exit 0), then the launch list, then one `--set full` capture of K2.
# Outputs under gpurun_out/ (scratch); condensed here into profiles/ by tools/ncu_summary.py / tools/launch_summary.py.
set -u
W="diff_drive_K1M_T100 steering_K4096_T50 full_body_K16384_T100 batched_1024robots_K1024_T50"
for w in $W; do
  python tools/profile_workload.py --workload $w > gpurun_out/r02_plain_$w.log 2>&1 || { echo "plain run failed: $w"; tail -5 gpurun_out/r02_plain_$w.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_$w.csv \
      python tools/profile_workload.py --workload $w > gpurun_out/r02_ncu_launch_$w.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:rollout_cost -s 32 -c 2 -f -o gpurun_out/r02_prof_k2_$w \
      python tools/profile_workload.py --workload $w > gpurun_out/r02_ncu_full_$w.log 2>&1
done
# the strong-scaling shard (K per GPU = 2^17) and the other kernels of the headline workload
python tools/profile_workload.py --workload diff_drive_K1M_T100 --K 131072 > gpurun_out/r02_plain_shard.log 2>&1 && {
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_shard_K131072.csv \
      python tools/profile_workload.py --workload diff_drive_K1M_T100 --K 131072 > gpurun_out/r02_ncu_launch_shard.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:rollout_cost -s 32 -c 2 -f -o gpurun_out/r02_prof_k2_shard_K131072 \
      python tools/profile_workload.py --workload diff_drive_K1M_T100 --K 131072 > gpurun_out/r02_ncu_full_shard.log 2>&1
}
ncu --set full --clock-control none --import-source on -k regex:"noise_kernel|rescale_tail|candidate_grid" -s 90 -c 3 -f -o gpurun_out/r02_prof_other_diff_drive_K1M_T100 \
    python tools/profile_workload.py --workload diff_drive_K1M_T100 > gpurun_out/r02_ncu_full_other.log 2>&1
ls -la gpurun_out/r02_* | head -40
