#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests -x -q -m gpu -k "ensi or utem or edge_cases or cxx" 2>&1 | tail -4
GPP_TRACE=1 timeout 600 python profiles/ensi_multi_time.py 2>&1 | grep -v "ensi_multi_host\|^{" | tail -8 | cut -c1-260
