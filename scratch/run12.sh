#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
python profiles/ensi_device_time.py 2500 2 2>&1 | tail -2
for v in c22 c18 c16; do
  GPP_B200_LIB=$PWD/scratch/lib_$v.so python profiles/ensi_device_time.py 2500 2 2>&1 | tail -2
  GPP_B200_LIB=$PWD/scratch/lib_$v.so timeout 600 python -m pytest tests -x -q -m gpu -k "ensi" 2>&1 | tail -2
done
