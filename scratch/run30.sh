#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python profiles/next_rows_time.py 2000 > gpurun_out/r2_next_rows_time.json 2> gpurun_out/r2_next_rows_time.err; echo rc=$?; cut -c1-230 gpurun_out/r2_next_rows_time.err | tail -16
