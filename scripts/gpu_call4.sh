#!/bin/bash
# GPU call 4: parity suite (TMA-box tests in their own process), deferred DIFF march A/B, multi-GPU per-rank tuning, ncu of DIFF batches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "not tma_box" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest.log
timeout 600 python -m pytest tests -m gpu -q -k "tma_box" > gpurun_out/pytest_box.log 2>&1; echo "pytest box rc=$?"
tail -4 gpurun_out/pytest_box.log
timeout 900 python scripts/tune.py --steps 8 --split \
    "default:" "nodefer:SDFRT_DEFER=0" "default_i1:inflight=1" "nodefer_i1:SDFRT_DEFER=0;inflight=1" \
    > gpurun_out/tune4a.jsonl 2> gpurun_out/tune4a.err; echo "tune a rc=$?"
cut -c1-700 gpurun_out/tune4a.jsonl
timeout 900 python scripts/tune.py --steps 8 \
    "p2:partition_of=2" "p2_b23:partition_of=2,SDFRT_BATCH_LOG2=23" "p2_b24:partition_of=2,SDFRT_BATCH_LOG2=24" \
    "p4:partition_of=4" "p4_b22:partition_of=4,SDFRT_BATCH_LOG2=22" "p4_b23:partition_of=4,SDFRT_BATCH_LOG2=23" "p4_b24:partition_of=4,SDFRT_BATCH_LOG2=24" \
    "p8:partition_of=8" "p8_b22:partition_of=8,SDFRT_BATCH_LOG2=22" "p8_b23:partition_of=8,SDFRT_BATCH_LOG2=23" "p8_b24:partition_of=8,SDFRT_BATCH_LOG2=24" \
    "p8_b23_s8:partition_of=8,SDFRT_BATCH_LOG2=23,SDFRT_SLOTS=8" "p8_b23_i3:partition_of=8,SDFRT_BATCH_LOG2=23;inflight=3" "p8_b23_i1:partition_of=8,SDFRT_BATCH_LOG2=23;inflight=1" \
    "p8_b23_nodefer:partition_of=8,SDFRT_BATCH_LOG2=23,SDFRT_DEFER=0" \
    "n1_b24:SDFRT_BATCH_LOG2=24" "n1_b23:SDFRT_BATCH_LOG2=23" "n1_s8:SDFRT_SLOTS=8" "n1_ml16:SDFRT_MIN_LIVE=16" "n1_ml28:SDFRT_MIN_LIVE=28" \
    > gpurun_out/tune4b.jsonl 2> gpurun_out/tune4b.err; echo "tune b rc=$?"
cut -c1-330 gpurun_out/tune4b.jsonl
timeout 300 python scripts/tune.py --steps 8 "box:SDFRT_BOX=10" "box_rows4:SDFRT_BOX=10,SDFRT_ROWS8=0" "box_vshadow:SDFRT_BOX=2" "box_nodefer:SDFRT_BOX=10,SDFRT_DEFER=0" > gpurun_out/tune4c.jsonl 2> gpurun_out/tune4c.err; echo "tune box rc=$?"
cut -c1-330 gpurun_out/tune4c.jsonl; tail -3 gpurun_out/tune4c.err
FLAGS="--steps 1 --warmup 0 --no-cpu-baseline --no-secondary --no-grad-check --no-e2e --views-in-flight 1"
summarise() {  # $1 = report base name
    [ -f gpurun_out/$1.ncu-rep ] || return
    python profiles/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1.md 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    sz=$(stat -c %s gpurun_out/$1.ncu-rep)
    if [ "$sz" -gt 9000000 ]; then rm -f gpurun_out/$1.ncu-rep; fi
}
K='k_march|k_shade|k_gen_primary|k_shadow_splat'
timeout 900 ncu --set full --clock-control none -k regex:"$K" -s 25 -c 5 -o gpurun_out/r2_c3_diff python bench.py $FLAGS > gpurun_out/ncu2.log 2>&1; echo "ncu diff rc=$?"
summarise r2_c3_diff
timeout 900 ncu --set full --clock-control none -k regex:k_march -s 24 -c 1 -o gpurun_out/r2_c3_deferred python bench.py $FLAGS > gpurun_out/ncu6.log 2>&1; echo "ncu deferred rc=$?"
summarise r2_c3_deferred
SDFRT_BOX=10 SDFRT_DEFER=0 timeout 900 ncu --set full --clock-control none -k regex:k_march -s 9 -c 1 -o gpurun_out/r2_c3_shadow_box python bench.py $FLAGS > gpurun_out/ncu3.log 2>&1; echo "ncu box rc=$?"
summarise r2_c3_shadow_box
SDFRT_DEFER=0 timeout 900 ncu --set full --clock-control none -k regex:k_march -s 9 -c 1 -o gpurun_out/r2_c3_shadow_rows8 python bench.py $FLAGS > gpurun_out/ncu4.log 2>&1; echo "ncu rows8 rc=$?"
summarise r2_c3_shadow_rows8
du -sh gpurun_out; ls gpurun_out
