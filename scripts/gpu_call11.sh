#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "prb or deferred" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 600 python scripts/prb_timing.py > gpurun_out/prb_timing_defer.log 2>&1; echo "prb rc=$?"; tail -30 gpurun_out/prb_timing_defer.log
SDFRT_DEFER=0 timeout 600 python scripts/prb_timing.py > gpurun_out/prb_timing_nodefer.log 2>&1; echo "prb rc=$?"; grep "_ms" gpurun_out/prb_timing_nodefer.log
timeout 600 python scripts/conv_timing.py > gpurun_out/conv_timing.log 2>&1; echo "conv rc=$?"; tail -8 gpurun_out/conv_timing.log
timeout 600 python bench.py --workload c5_prb_blobs256_512x512_128spp --steps 3 --warmup 1 --no-secondary --no-e2e --no-grad-check > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"; tail -2 gpurun_out/bench_c5.err
