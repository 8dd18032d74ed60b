#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/tc_probe.py > gpurun_out/probe.log 2>&1
grep -E "TC_PROBE|max err|first bad|bad rows|got|exp" gpurun_out/probe.log | head -80
REFVSR_TC_LAYOUT=0 timeout -s KILL 200 python tools/conv_sweep.py > gpurun_out/sweep0.log 2>&1
timeout -s KILL 200 python tools/conv_sweep.py > gpurun_out/sweep1.log 2>&1
echo "--- layout0"; cat gpurun_out/sweep0.log; echo "--- layout auto"; cat gpurun_out/sweep1.log
