#!/bin/bash
# GPU call 1 of round 2: parity of the rewritten OI kernel / EnSI device entry, variant timings, ncu of the OI kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest1.log
tail -5 $O/r2_pytest1.log
{
for rows in 4000 500; do
  python profiles/oi_time.py $rows
  for v in top3 lru16 stats; do GPP_B200_LIB=$PWD/scratch/lib_$v.so python profiles/oi_time.py $rows; done
done
python profiles/ensi_device_time.py 2500 2
GPP_B200_LIB=$PWD/scratch/lib_regj.so python profiles/ensi_device_time.py 2500 2
python profiles/oi_general_time.py 1000
} > $O/r2_time1.log 2>&1
cat $O/r2_time1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oi_fast_kernel -s 2 -c 1 -o $O/r2_oi_fast_v1 -f python profiles/oi_probe.py fast > $O/r2_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oi_chol_kernel -s 2 -c 1 -o $O/r2_oi_chol_v1 -f python profiles/oi_probe.py chol >> $O/r2_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oi_general_kernel -s 2 -c 1 -o $O/r2_oi_general_v1 -f python profiles/oi_probe.py general >> $O/r2_ncu1.log 2>&1
tail -3 $O/r2_ncu1.log
ls -la $O | tail -8
