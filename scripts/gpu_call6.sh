#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/tma_test scripts/tma_test.cu > gpurun_out/tma_build.log 2>&1
/tmp/tma_test > gpurun_out/tma_test.log 2>&1; echo "tma_test rc=$?"; cat gpurun_out/tma_test.log
timeout 600 python -m pytest tests -m gpu -q -k "tma_box" > gpurun_out/pytest_box.log 2>&1; echo "pytest box rc=$?"; tail -4 gpurun_out/pytest_box.log
timeout 400 python scripts/tune.py --steps 8 --split "nodefer:SDFRT_DEFER=0" "box_nodefer:SDFRT_BOX=10,SDFRT_DEFER=0" "box_vshadow:SDFRT_BOX=2" "box_rows4:SDFRT_BOX=10,SDFRT_ROWS8=0,SDFRT_DEFER=0" "rows4_nodefer:SDFRT_ROWS8=0,SDFRT_DEFER=0" > gpurun_out/tune6.jsonl 2> gpurun_out/tune6.err; echo "tune rc=$?"
cut -c1-900 gpurun_out/tune6.jsonl; tail -3 gpurun_out/tune6.err
FLAGS="--steps 1 --warmup 0 --no-cpu-baseline --no-secondary --no-grad-check --no-e2e --views-in-flight 1"
summarise() {
    [ -f gpurun_out/$1.ncu-rep ] || return
    python profiles/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1.md 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    sz=$(stat -c %s gpurun_out/$1.ncu-rep); if [ "$sz" -gt 9000000 ]; then rm -f gpurun_out/$1.ncu-rep; fi
}
SDFRT_BOX=10 SDFRT_DEFER=0 timeout 900 ncu --set full --clock-control none -k regex:k_march -s 9 -c 1 -o gpurun_out/r2_c3_shadow_box python bench.py $FLAGS > gpurun_out/ncu3.log 2>&1; echo "ncu box rc=$?"
summarise r2_c3_shadow_box
ls -la gpurun_out
