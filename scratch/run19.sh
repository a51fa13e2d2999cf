#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python profiles/oi_general_time.py 1000 2>&1 | tail -6
timeout 900 python -m pytest tests -x -q -m gpu -k "oi or optimal or structure or ensi_multi or spatial or cxx" 2>&1 | tail -4
