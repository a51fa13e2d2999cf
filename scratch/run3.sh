#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python scratch/debug_std.py > $O/r2_debug_std.log 2>&1; cat $O/r2_debug_std.log
timeout 1800 python -m pytest tests -m gpu -q > $O/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest3.log
tail -25 $O/r2_pytest3.log | cut -c1-400
{
python profiles/nbh_time.py 4000 --qf
python profiles/nbh_time.py 8000 --qf
python profiles/oi_time.py 4000
python profiles/oi_time.py 500
} > $O/r2_time3.log 2>&1
cat $O/r2_time3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbh_sumf_tma_kernel -s 2 -c 1 -o $O/r2_nbh_mean_v2 -f python profiles/nbh_probe.py mean > $O/r2_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qf_tma_kernel -s 2 -c 1 -o $O/r2_qf_v2 -f python profiles/nbh_probe.py qf >> $O/r2_ncu3.log 2>&1
tail -3 $O/r2_ncu3.log
