#!/bin/bash
# GPU call: parity tests, bench lines of both arms, A/B tuning runs, ncu captures reduced to CSV/markdown on the box
# (gpurun_out/ must stay below 64 MiB: .ncu-rep files are summarised here and only kept when small)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python scripts/tune.py --split \
    "base1:inflight=1" "base2:inflight=2" "base3:inflight=3" \
    "r8_vshadow:SDFRT_ROWS8=2" "r8_shadow:SDFRT_ROWS8=10" "r8_all:SDFRT_ROWS8=15" \
    "p8_1:partition_of=8;inflight=1" "p8_2:partition_of=8;inflight=2" "p8_4:partition_of=8;inflight=4" \
    "p8_2_s8:partition_of=8;inflight=2,SDFRT_SLOTS=8" "p8_2_b21:partition_of=8;inflight=2,SDFRT_BATCH_LOG2=21" \
    "p8_2_b20:partition_of=8;inflight=2,SDFRT_BATCH_LOG2=20" "p8_2_b22:partition_of=8;inflight=2,SDFRT_BATCH_LOG2=22" \
    > gpurun_out/tune.jsonl 2> gpurun_out/tune.err; echo "tune rc=$?"
cat gpurun_out/tune.jsonl | cut -c1-400
FLAGS="--steps 1 --warmup 0 --no-cpu-baseline --no-secondary --no-grad-check --no-e2e --views-in-flight 1"
summarise() {  # $1 = report base name
    python profiles/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1.md 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    sz=$(stat -c %s gpurun_out/$1.ncu-rep)
    if [ "$sz" -gt 9000000 ]; then rm -f gpurun_out/$1.ncu-rep; fi
}
timeout 1200 ncu --set full --clock-control none -k regex:'k_march|k_shade|k_gen_primary|k_shadow_splat' -c 12 \
    -o gpurun_out/r2_c3_forward python bench.py $FLAGS > gpurun_out/ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
summarise r2_c3_forward
timeout 600 ncu --set full --clock-control none -k regex:'k_backward|k_film_adjoint' -c 2 \
    -o gpurun_out/r2_c3_backward python bench.py $FLAGS > gpurun_out/ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
summarise r2_c3_backward
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 500 --csv \
    --log-file gpurun_out/r2_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-grad-check --no-e2e --views-in-flight 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'k_prb|k_march' -c 12 \
    -o gpurun_out/r2_c5_prb python bench.py --workload c5_prb_blobs256_512x512_128spp $FLAGS > gpurun_out/ncu_prb.log 2>&1; echo "ncu prb rc=$?"
summarise r2_c5_prb
du -sh gpurun_out; ls -la gpurun_out
