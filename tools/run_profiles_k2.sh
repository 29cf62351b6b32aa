#!/bin/bash
# K2 changed (end phase): refresh its launch share and `--set full` counters for the headline workload and the shard.
set -u
w=diff_drive_K1M_T100
python tools/profile_workload.py --workload $w > gpurun_out/r02_plain_$w.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_$w.csv \
    python tools/profile_workload.py --workload $w > gpurun_out/r02_ncu_launch_$w.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_cost -s 32 -c 2 -f -o gpurun_out/r02_prof_k2_$w \
    python tools/profile_workload.py --workload $w > gpurun_out/r02_ncu_full_$w.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_shard_K131072.csv \
    python tools/profile_workload.py --workload $w --K 131072 > gpurun_out/r02_ncu_launch_shard.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_cost -s 32 -c 2 -f -o gpurun_out/r02_prof_k2_shard_K131072 \
    python tools/profile_workload.py --workload $w --K 131072 > gpurun_out/r02_ncu_full_shard.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_batched_1024robots_K1024_T50.csv \
    python tools/profile_workload.py --workload batched_1024robots_K1024_T50 > gpurun_out/r02_ncu_launch_batched.log 2>&1
ls -la gpurun_out/r02_prof_k2_* gpurun_out/r02_launches_*
