#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
( time timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_ref_v2.json 2> gpurun_out/r2_bench_ref_v2.err ) 2>&1 | grep real
cut -c1-600 gpurun_out/r2_bench_ref_v2.json; tail -2 gpurun_out/r2_bench_ref_v2.err | cut -c1-200
