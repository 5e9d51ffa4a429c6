#!/bin/bash
# Multi-GPU pass on one box:  bash tools/gpu_n.sh <N>   -- 2-rank NCCL test, then bench.py at N GPUs (torchrun, as the driver does)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
if [ "$N" == "2" ]; then
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_n2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_n2.log
fi
echo "== bench N=$N"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-extra > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cut -c1-2600 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
