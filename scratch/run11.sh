#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $O/r2_pytest7.log
timeout 900 python bench.py > $O/r2_bench_v2.json 2> $O/r2_bench_v2.err; echo "bench rc=$?"; tail -3 $O/r2_bench_v2.err | cut -c1-300
python -c "
import json
d=json.load(open('$O/r2_bench_v2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e'], d['roofline_fp64']['frac'])
"
