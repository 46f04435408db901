#!/bin/bash
# 2-GPU call: NCCL parity test, the bench line at N = 2 exactly as the driver launches it, the reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"
cut -c1-600 gpurun_out/bench_n2.json; cut -c1-500 gpurun_out/bench_ref_n2.json
