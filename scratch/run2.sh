#!/bin/bash
# GPU call 2 of round 2: float mean kernel, bracket-tracking quantile_fast, statistics family, reference suite
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest2.log
tail -15 $O/r2_pytest2.log
{
python profiles/nbh_time.py 4000 --qf
GPP_NBH_F64=1 python profiles/nbh_time.py 4000
python profiles/nbh_time.py 8000 --qf
} > $O/r2_time2.log 2>&1
cat $O/r2_time2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbh_sumf_tma_kernel -s 2 -c 1 -o $O/r2_nbh_mean_v1 -f python profiles/nbh_probe.py mean > $O/r2_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qf_tma_kernel -s 2 -c 1 -o $O/r2_qf_v1 -f python profiles/nbh_probe.py qf >> $O/r2_ncu2.log 2>&1
tail -3 $O/r2_ncu2.log
ls -la $O | tail -6
