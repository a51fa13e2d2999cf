#!/bin/bash
# 8-GPU evidence: worker parity checks, the N=8 and N=4 bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/r2_gpus8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_worker.py > $O/r2_mgpu_worker_n8_v3.json 2> $O/r2_mgpu_worker_n8_v3.err; echo "worker rc=$?"; cut -c1-2500 $O/r2_mgpu_worker_n8_v3.json; tail -5 $O/r2_mgpu_worker_n8_v3.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2_bench_n8_v3.json 2> $O/r2_bench_n8_v3.err; echo "bench8 rc=$?"; tail -5 $O/r2_bench_n8_v3.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r2_bench_n4_v3.json 2> $O/r2_bench_n4_v3.err; echo "bench4 rc=$?"; tail -5 $O/r2_bench_n4_v3.err | cut -c1-300
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.load(open('gpurun_out/r2_bench_n%d_v1.json' % n))
    except Exception as e:
        print(n, 'no json', e); continue
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'n_gpus')}, d['e2e'], d['config'].get('sharded_equals_whole'))
    for k, v in d.get('secondary_multi_gpu', {}).items(): print(' ', k, v)
PY
