#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/r2_gpus2.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $O/r2_pytest_multi_v3.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest_multi_v3.log; tail -30 $O/r2_pytest_multi_v3.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_worker.py > $O/r2_mgpu_worker_n2_v3.json 2> $O/r2_mgpu_worker_n2_v3.err; echo "worker rc=$?"; cat $O/r2_mgpu_worker_n2_v3.json | cut -c1-3000; tail -5 $O/r2_mgpu_worker_n2_v3.err | cut -c1-300
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench_n2_v3.json 2> $O/r2_bench_n2_v3.err; echo "bench rc=$?"; tail -5 $O/r2_bench_n2_v3.err | cut -c1-300
python -c "
import json
d=json.load(open('$O/r2_bench_n2_v3.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e'], d['config'])
for k,v in d.get('secondary_multi_gpu',{}).items(): print(k, v)
"
