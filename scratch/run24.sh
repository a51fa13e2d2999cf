#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $O/r2_pytest12.log
( time timeout 900 python bench.py > $O/r2_bench_v6.json 2> $O/r2_bench_v6.err ) 2>&1 | grep real; tail -2 $O/r2_bench_v6.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v6.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e'], d['roofline_fp64']['frac'], d['clocks'])
for k,v in d['secondary'].items():
    if isinstance(v, dict): print(k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms','frac_of_hbm_peak','kernel_ms','seconds_end_to_end','gridpoints/s')})
PY
