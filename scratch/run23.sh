#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
# 400 rows of C5 (1 M points): enough for steady state, short under ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ensi_kernel -s 1 -c 1 -o $O/r2_ensi_v1 -f python profiles/ensi_device_time.py 400 1 > $O/r2_ncu7.log 2>&1; tail -2 $O/r2_ncu7.log
