#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python bench.py > $O/r2_bench_v3.json 2> $O/r2_bench_v3.err; echo "bench rc=$?"; tail -3 $O/r2_bench_v3.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v3.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e'], d['roofline_fp64']['frac'])
for k in ('ensi','ensi_multi_ebesc'):
    print(k, json.dumps(d['secondary'].get(k))[:1500])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oi_fast_kernel -s 2 -c 1 -o $O/r2_oi_fast_v2 -f python profiles/oi_probe.py fast > $O/r2_ncu5.log 2>&1; tail -2 $O/r2_ncu5.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_v1.csv python bench.py --steps 2 --warmup 1 --quick > $O/r2_launch_bench.log 2>&1; echo "launch list rc=$?"; wc -l $O/r2_launches_v1.csv
