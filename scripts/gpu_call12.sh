#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/prb_timing.py > gpurun_out/prb_timing_defer.log 2>&1; echo "prb rc=$?"; tail -12 gpurun_out/prb_timing_defer.log
SDFRT_PRB_OCC=3 timeout 600 python scripts/prb_timing.py > gpurun_out/prb_timing_occ3.log 2>&1; echo "prb rc=$?"; grep "_ms" gpurun_out/prb_timing_occ3.log
