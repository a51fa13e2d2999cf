#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oi_general_kernel -s 2 -c 1 -o $O/r2_oi_general_v2 -f python profiles/oi_probe.py general > $O/r2_ncu6.log 2>&1; tail -2 $O/r2_ncu6.log
