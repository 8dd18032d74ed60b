#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc --launch-skip 2 --launch-count 1 -f -o gpurun_out/r01_conv_lr python tools/ncu_conv.py > gpurun_out/ncu_conv_lr.log 2>&1
tail -2 gpurun_out/ncu_conv_lr.log
timeout -s KILL 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_window.csv python tools/profile_window.py > gpurun_out/ncu_window.log 2>&1
tail -2 gpurun_out/ncu_window.log; wc -l gpurun_out/r01_launches_window.csv
ls -la gpurun_out/*.ncu-rep
