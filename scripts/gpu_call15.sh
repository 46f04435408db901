#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "not tma_box" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
timeout 600 python scripts/prb_timing.py > gpurun_out/prb_timing.log 2>&1; echo "prb rc=$?"; grep "_ms" gpurun_out/prb_timing.log | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5_prb_blobs256_512x512_128spp --steps 1 --warmup 1 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_prb_backward|k_scatter_sites|k_prb_flag|k_prb_wf_emit|k_prb_primal" -s 12 -c 6 -o gpurun_out/c5_prb_r2 python bench.py --workload c5_prb_blobs256_512x512_128spp --steps 1 --warmup 1 --no-secondary --no-e2e --no-grad-check --no-cpu-baseline > gpurun_out/ncu_c5_full.log 2>&1; echo "ncu full rc=$?"
python profiles/ncu_summary.py gpurun_out/c5_prb_r2.ncu-rep > gpurun_out/r2_c5_prb_queue.md 2> gpurun_out/ncu_summary.err; echo "summary rc=$?"
ls -la gpurun_out/*.ncu-rep; rm -f gpurun_out/*.ncu-rep
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
ls -la gpurun_out
