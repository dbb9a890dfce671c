#!/bin/bash
# GPU evidence pass for profiles/: closed-loop step breakdown (config 4), launch list of the closed
# loop, and ncu --set full captures of the solve kernel at B = 1 (latency) and at N = 15 (production
# horizon).  Every step runs once WITHOUT ncu first (a number printed under ncu is never reported).
#   gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh'
set -u
O=gpurun_out
mkdir -p $O
T="timeout 280"
$T python tools/rollout_rate.py 16384 40 t4 > $O/cfg4_step_breakdown.txt 2>&1
echo "rc rollout_rate $?" >> $O/cfg4_step_breakdown.txt
$T python tools/prof_one.py 1 warp 30 3 1 > $O/prof_b1.txt 2>&1
$T python tools/prof_one.py 16384 warp 15 3 1 > $O/prof_n15.txt 2>&1
# launch list of the closed loop (ungraphed ticks + graph replays)
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/cfg4_launches.csv \
    python tools/rollout_rate.py 16384 8 t4 > $O/cfg4_ncu.log 2>&1
echo "rc launches $?" >> $O/cfg4_ncu.log
$T ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 2 -c 1 -f -o $O/r2_B1 \
    python tools/prof_one.py 1 warp 30 3 1 > $O/ncu_b1.log 2>&1
echo "rc b1 $?" >> $O/ncu_b1.log
$T ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 2 -c 1 -f -o $O/r2_N15 \
    python tools/prof_one.py 16384 warp 15 3 1 > $O/ncu_n15.log 2>&1
echo "rc n15 $?" >> $O/ncu_n15.log
ls -la $O | tail -n 12
