#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests -x -q -m gpu -k "ensi_multi or cxx" 2>&1 | tail -6
bash scratch/run21.sh 2>&1 | grep -v "^\[gpp trace\].*observation" | tail -10
