#!/bin/bash
# GPU call: parity tests, the bench line of both arms, ncu captures of the shipped kernels (round 2 record)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
FLAGS="--steps 1 --warmup 0 --no-cpu-baseline --no-secondary --no-grad-check --no-e2e"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_march|k_shade|k_gen_primary|k_shadow_splat' -c 22 \
    -o gpurun_out/r2_c3_forward python bench.py $FLAGS > gpurun_out/ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_backward|k_film_adjoint' -c 2 \
    -o gpurun_out/r2_c3_backward python bench.py $FLAGS > gpurun_out/ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 500 --csv \
    --log-file gpurun_out/r2_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-grad-check --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_prb|k_march' -c 16 \
    -o gpurun_out/r2_c5_prb python bench.py --workload c5_prb_blobs256_512x512_128spp $FLAGS > gpurun_out/ncu_prb.log 2>&1; echo "ncu prb rc=$?"
ls -la gpurun_out
