#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/nmma_sweep.log
timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv or resblock" > gpurun_out/nmma_tests.log 2>&1
tail -3 gpurun_out/nmma_tests.log
for L in 0 1; do for N in 1 3; do
  echo "== layout $L nmma<=$N" >> gpurun_out/nmma_sweep.log
  REFVSR_TC_LAYOUT=$L REFVSR_NMMA=$N timeout -s KILL 60 python tools/conv_sweep.py >> gpurun_out/nmma_sweep.log 2>&1
done; done
REFVSR_TC_LAYOUT=1 timeout -s KILL 200 python tools/conv_hist.py > gpurun_out/conv_hist.log 2>&1
timeout -s KILL 200 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/nmma_bench.json 2> gpurun_out/nmma_bench.err
REFVSR_TC_LAYOUT=1 timeout -s KILL 200 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/nmma_bench_l1.json 2>> gpurun_out/nmma_bench.err
REFVSR_TC_LAYOUT=1 REFVSR_NMMA=1 timeout -s KILL 200 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/nmma_bench_l1n1.json 2>> gpurun_out/nmma_bench.err
python - <<'PY'
import json
for f in ('nmma_bench','nmma_bench_l1','nmma_bench_l1n1'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
