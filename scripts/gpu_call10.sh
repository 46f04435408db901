#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "convolution or envmap or compat" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 600 python scripts/conv_timing.py > gpurun_out/conv_timing.log 2>&1; echo "conv rc=$?"; tail -8 gpurun_out/conv_timing.log
timeout 600 python bench.py --workload c5_prb_blobs256_512x512_128spp --steps 3 --warmup 1 --no-secondary --no-e2e --no-grad-check > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"; tail -2 gpurun_out/bench_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5_prb_blobs256_512x512_128spp --steps 1 --warmup 1 --no-secondary --no-e2e --no-grad-check > gpurun_out/ncu_c5.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
