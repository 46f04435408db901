#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q -x -k "prb and golden" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
timeout 100 python bench.py --steps 3 --warmup 3 --no-secondary --no-grad-check --no-cpu-baseline --no-e2e > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_short.json
