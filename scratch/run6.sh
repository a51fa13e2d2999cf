#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_reference_suite.py -m gpu -q > $O/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest6.log; tail -3 $O/r2_pytest6.log
timeout 1500 python bench.py --steps 10 --warmup 3 > $O/r2_bench_v1.json 2> $O/r2_bench_v1.err; echo "bench rc=$?"; tail -5 $O/r2_bench_v1.err
python -c "
import json
d=json.load(open('$O/r2_bench_v1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline'], d.get('roofline_fp64'), d.get('cpu_baseline'))
for k,v in d.get('secondary',{}).items(): print(k, v)
"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_ref_v1.json 2> $O/r2_bench_ref_v1.err; echo "ref rc=$?"; cat $O/r2_bench_ref_v1.json | cut -c1-1500
