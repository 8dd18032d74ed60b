#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/tc_probe.py > gpurun_out/probe.log 2>&1
timeout -s KILL 300 python tools/profile_kernels.py > gpurun_out/kern.json 2> gpurun_out/kern.err
grep TC_PROBE gpurun_out/probe.log; grep -E '"seconds"|"frac"' gpurun_out/kern.json
