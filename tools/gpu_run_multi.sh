#!/bin/bash
# dev tool (GPU box with N GPUs): the library's NCCL slab / batch modes against one GPU, then the bench at N
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
T=${2:-r2m}
nvidia-smi -L | head -8
mkdir -p /tmp/mg && rm -f /tmp/mg/ok
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_worker.py /tmp/mg > gpurun_out/multi_worker_${T}_n$N.log 2>&1
echo "worker rc=$? ok-file: $(ls /tmp/mg)"; tail -5 gpurun_out/multi_worker_${T}_n$N.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${T}_n$N.json 2> gpurun_out/bench_${T}_n$N.err
tail -c 1800 gpurun_out/bench_${T}_n$N.json; echo; tail -5 gpurun_out/bench_${T}_n$N.err
