#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for rows in 500 1000; do
  python profiles/e2e_probe.py $rows | tail -1
  for ch in 2 4 8; do GPP_OI_PIPELINE_MIN=100000 GPP_OI_CHUNKS=$ch python profiles/e2e_probe.py $rows | tail -1; done
done
GPP_TRACE=1 python profiles/e2e_probe.py 500 2>&1 | tail -12
