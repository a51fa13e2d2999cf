#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests -x -q -m gpu -k "ensi_multi or cxx or edge_cases" 2>&1 | tail -3
bash scratch/run21.sh 2>&1 | grep -v "^\[gpp trace\].*observation" | tail -8
timeout 300 python -c "
import json, bench, gridpp_b200 as gpp
m = bench.ensi_multi_metric(gpp); print(m['seconds_end_to_end'], m['cpu_baseline']['gpu_vs_cpu_max_rel_err'])
"
