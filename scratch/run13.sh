#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
GPP_B200_LIB=$PWD/scratch/lib_qfstats.so python profiles/qf_stats.py 4000 2>&1 | tail -8
