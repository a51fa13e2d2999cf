#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
GPP_B200_LIB=$PWD/scratch/lib_ydouble.so python profiles/ensi_device_time.py 2500 2 2>&1 | tail -1
python profiles/ensi_device_time.py 2500 2 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu -k "ensi or utem" 2>&1 | tail -2
