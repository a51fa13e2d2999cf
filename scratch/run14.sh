#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python profiles/sanitize_all.py 2>&1 | tail -3
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python profiles/sanitize_all.py > $O/r2_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/r2_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python profiles/sanitize_all.py > $O/r2_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "Race reported\|hazard" $O/r2_racecheck.log; tail -4 $O/r2_racecheck.log
