#!/bin/bash
# Final round-2 refresh of the ncu evidence after the tail / PDL changes (run on the GPU box):
#   launch lists of every workload + the strong-scaling shard, `--set full` of K2 for the headline workload and the
#   shard, and of the other kernels (generator, tail, candidate grid) of the headline workload.
set -u
W="diff_drive_K1M_T100 steering_K4096_T50 full_body_K16384_T100 batched_1024robots_K1024_T50"
for w in $W; do
  python tools/profile_workload.py --workload $w > gpurun_out/r02_plain_$w.log 2>&1 || { echo "plain run failed: $w"; tail -5 gpurun_out/r02_plain_$w.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_$w.csv \
      python tools/profile_workload.py --workload $w > gpurun_out/r02_ncu_launch_$w.log 2>&1
done
w=diff_drive_K1M_T100
ncu --set full --clock-control none --import-source on -k regex:rollout_cost -s 32 -c 2 -f -o gpurun_out/r02_prof_k2_$w \
    python tools/profile_workload.py --workload $w > gpurun_out/r02_ncu_full_$w.log 2>&1
python tools/profile_workload.py --workload $w --K 131072 > gpurun_out/r02_plain_shard.log 2>&1 && {
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_shard_K131072.csv \
      python tools/profile_workload.py --workload $w --K 131072 > gpurun_out/r02_ncu_launch_shard.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:rollout_cost -s 32 -c 2 -f -o gpurun_out/r02_prof_k2_shard_K131072 \
      python tools/profile_workload.py --workload $w --K 131072 > gpurun_out/r02_ncu_full_shard.log 2>&1
}
ncu --set full --clock-control none --import-source on -k regex:"noise_kernel|rescale_tail|candidate_grid" -s 90 -c 3 -f -o gpurun_out/r02_prof_other_$w \
    python tools/profile_workload.py --workload $w > gpurun_out/r02_ncu_full_other.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rescale_tail" -s 30 -c 1 -f -o gpurun_out/r02_prof_tail_shard_K131072 \
    python tools/profile_workload.py --workload $w --K 131072 > gpurun_out/r02_ncu_full_tail_shard.log 2>&1
ls -la gpurun_out/r02_prof_* gpurun_out/r02_launches_* | head -40
