#!/bin/bash
# GPU call 3: full parity suite, tuning of batch size / slots / views in flight for the multi-GPU per-rank workloads, TMA box A/B,
# ncu captures of representative launches (skipping the empty top-of-film batches), host CPU facts for the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
(nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu | head -25; free -g | head -2) > gpurun_out/host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest.log
timeout 900 python scripts/tune.py --steps 8 \
    "default:" "rows4:SDFRT_ROWS8=0" "box:SDFRT_BOX=10" "box_rows4:SDFRT_BOX=10,SDFRT_ROWS8=0" "box_vshadow:SDFRT_BOX=2" \
    "p2:partition_of=2" "p2_b23:partition_of=2,SDFRT_BATCH_LOG2=23" "p2_b24:partition_of=2,SDFRT_BATCH_LOG2=24" "p2_b24_s8:partition_of=2,SDFRT_BATCH_LOG2=24,SDFRT_SLOTS=8" \
    "p4:partition_of=4" "p4_b22:partition_of=4,SDFRT_BATCH_LOG2=22" "p4_b23:partition_of=4,SDFRT_BATCH_LOG2=23" "p4_b24:partition_of=4,SDFRT_BATCH_LOG2=24" "p4_b23_s8:partition_of=4,SDFRT_BATCH_LOG2=23,SDFRT_SLOTS=8" \
    "p8:partition_of=8" "p8_b22:partition_of=8,SDFRT_BATCH_LOG2=22" "p8_b23:partition_of=8,SDFRT_BATCH_LOG2=23" "p8_b24:partition_of=8,SDFRT_BATCH_LOG2=24" \
    "p8_b23_s8:partition_of=8,SDFRT_BATCH_LOG2=23,SDFRT_SLOTS=8" "p8_b23_i3:partition_of=8,SDFRT_BATCH_LOG2=23;inflight=3" "p8_b23_i1:partition_of=8,SDFRT_BATCH_LOG2=23;inflight=1" \
    "n1_b24:SDFRT_BATCH_LOG2=24" "n1_b23:SDFRT_BATCH_LOG2=23" "n1_s8:SDFRT_SLOTS=8" "n1_ml16:SDFRT_MIN_LIVE=16" "n1_ml28:SDFRT_MIN_LIVE=28" \
    > gpurun_out/tune3.jsonl 2> gpurun_out/tune3.err; echo "tune rc=$?"
cut -c1-330 gpurun_out/tune3.jsonl
timeout 300 python scripts/tune.py --steps 3 --warmup 1 --workload c5_prb_blobs256_512x512_128spp "c5:" "c5_b23:SDFRT_BATCH_LOG2=23" "c5_i1:inflight=1" > gpurun_out/tune3_c5.jsonl 2>> gpurun_out/tune3.err
cut -c1-330 gpurun_out/tune3_c5.jsonl
FLAGS="--steps 1 --warmup 0 --no-cpu-baseline --no-secondary --no-grad-check --no-e2e --views-in-flight 1"
summarise() {  # $1 = report base name
    python profiles/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1.md 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    sz=$(stat -c %s gpurun_out/$1.ncu-rep)
    if [ "$sz" -gt 9000000 ]; then rm -f gpurun_out/$1.ncu-rep; fi
}
K='k_march|k_shade|k_gen_primary|k_shadow_splat'
timeout 900 ncu --set full --clock-control none -k regex:"$K" -s 20 -c 5 -o gpurun_out/r2_c3_value python bench.py $FLAGS > gpurun_out/ncu1.log 2>&1; echo "ncu value rc=$?"
summarise r2_c3_value
timeout 900 ncu --set full --clock-control none -k regex:"$K" -s 5 -c 5 -o gpurun_out/r2_c3_diff python bench.py $FLAGS > gpurun_out/ncu2.log 2>&1; echo "ncu diff rc=$?"
summarise r2_c3_diff
SDFRT_BOX=10 timeout 900 ncu --set full --clock-control none -k regex:k_march -s 9 -c 1 -o gpurun_out/r2_c3_shadow_box python bench.py $FLAGS > gpurun_out/ncu3.log 2>&1; echo "ncu box rc=$?"
summarise r2_c3_shadow_box
SDFRT_ROWS8=0 timeout 900 ncu --set full --clock-control none -k regex:k_march -s 9 -c 1 -o gpurun_out/r2_c3_shadow_rows4 python bench.py $FLAGS > gpurun_out/ncu4.log 2>&1; echo "ncu rows4 rc=$?"
summarise r2_c3_shadow_rows4
timeout 600 ncu --set full --clock-control none -k regex:'k_backward' -c 1 -o gpurun_out/r2_c3_backward python bench.py $FLAGS > gpurun_out/ncu5.log 2>&1; echo "ncu bwd rc=$?"
summarise r2_c3_backward
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -c 500 --csv \
    --log-file gpurun_out/r2_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-grad-check --no-e2e --views-in-flight 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'k_prb|k_march|k_gen_primary' -s 44 -c 11 \
    -o gpurun_out/r2_c5_prb python bench.py --workload c5_prb_blobs256_512x512_128spp $FLAGS > gpurun_out/ncu_prb.log 2>&1; echo "ncu prb rc=$?"
summarise r2_c5_prb
timeout 600 ncu --set full --clock-control none -k regex:'k_prb_backward' -c 1 \
    -o gpurun_out/r2_c5_prb_bwd python bench.py --workload c5_prb_blobs256_512x512_128spp $FLAGS > gpurun_out/ncu_prb2.log 2>&1; echo "ncu prb bwd rc=$?"
summarise r2_c5_prb_bwd
du -sh gpurun_out; ls gpurun_out
