#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
python profiles/ensi_device_time.py 2500 2 2>&1 | tail -1
timeout 900 python -m pytest tests -x -q -m gpu -k "ensi or cxx or utem" 2>&1 | tail -2
