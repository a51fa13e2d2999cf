#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oi_chol_kernel -s 2 -c 1 -o $O/r2_oi_chol_v2 -f python profiles/oi_probe.py chol > $O/r2_ncu8.log 2>&1; tail -1 $O/r2_ncu8.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ensi_multi_kernel -s 8 -c 1 -o $O/r2_ensi_multi_v1 -f python profiles/ensi_multi_probe.py > $O/r2_ncu9.log 2>&1; tail -1 $O/r2_ncu9.log
