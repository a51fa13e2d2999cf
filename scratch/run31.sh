#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
GPP_TRACE=1 timeout 900 python profiles/ensi_multi_time.py > gpurun_out/r2_ensi_multi_time.json 2> gpurun_out/r2_ensi_multi_time.err; echo rc=$?; cut -c1-300 gpurun_out/r2_ensi_multi_time.err | grep -v "ensi_multi_host" | tail -12
