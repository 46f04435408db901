#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/tma_test scripts/tma_test.cu > gpurun_out/tma_build.log 2>&1
/tmp/tma_test > gpurun_out/tma_test.log 2>&1; echo "tma_test rc=$?"; cat gpurun_out/tma_test.log
timeout 300 compute-sanitizer --tool memcheck /tmp/tma_test > gpurun_out/tma_sanitizer.log 2>&1; tail -15 gpurun_out/tma_sanitizer.log
SDFRT_BOX=10 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/tune.py --workload c2_blobs128_256x256_16spp --steps 1 --warmup 0 "box:SDFRT_BOX=10;inflight=1" > gpurun_out/box_sanitizer.log 2>&1; grep -v "^$" gpurun_out/box_sanitizer.log | head -40
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --partition-of 8 --no-cpu-baseline --no-secondary --no-grad-check --no-e2e > gpurun_out/bench_p8.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --partition-of 4 --no-cpu-baseline --no-secondary --no-grad-check --no-e2e > gpurun_out/bench_p4.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --partition-of 2 --no-cpu-baseline --no-secondary --no-grad-check --no-e2e > gpurun_out/bench_p2.json 2>> gpurun_out/bench.err
ls -la gpurun_out
