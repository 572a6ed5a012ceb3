#!/bin/bash
# ncu captures of the bench's dominant kernel (rollout_kernel, 65,536 envs, 20 steps per launch); run under gpurun
# AFTER the same command has exited 0 without ncu.  Outputs under gpurun_out/.
set -e
TAG=${1:-r02}
python tools/rollout_one.py 65536 20 > gpurun_out/${TAG}_rollout_one.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/rollout_one.py 65536 20 > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o gpurun_out/${TAG}_rollout -f \
    python tools/rollout_one.py 65536 20 > gpurun_out/${TAG}_ncu_full.log 2>&1
# the same for the policy-fused variant (16-step launches of tools/policy_one.py)
python tools/policy_one.py > gpurun_out/${TAG}_policy_one.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/${TAG}_policy_rollout -f \
    python tools/policy_one.py > gpurun_out/${TAG}_ncu_policy.log 2>&1
