#!/bin/bash
# ncu evidence of round 2, one gpurun call:  gpurun --timeout 1700 -- 'bash tools/profile_round2.sh'
# (numbers printed by a run under ncu are never bench values)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# 1. launch list of the N = 1 bench step (configs[1] part only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_10m.csv \
    python bench.py --steps 2 --warmup 1 --no-cli --no-configs --no-twin --no-cpu --no-e2e --batch-meshes 0 > gpurun_out/r02_launches_bench.log 2>&1
# 2. full sections of the kernels of one configs[1] step (10M vertices)
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_decode_vertex_scan|k_encode_vtx_packed|k_vertex_candidates|k_bounds_reduce|k_requant|k_flatten|k_nocomp|k_scan_prep|k_vertex_order|k_gather_packed' \
    -c 40 -o gpurun_out/r02_kernels_10m python tools/profile_driver.py --nr 2237 --ns 4472 --reps 1 > gpurun_out/r02_kernels_10m.log 2>&1
# 3. the batch: launch list and the scan decoder over 195 meshes
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_batch195.csv \
    python tools/batch_probe.py --meshes 195 --reps 1 > gpurun_out/r02_launches_batch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_decode_vertex_scan|k_vertex_candidates_stage' -c 2 \
    -o gpurun_out/r02_batch195 python tools/batch_probe.py --meshes 195 --reps 1 > gpurun_out/r02_batch195.log 2>&1
ls -la gpurun_out | tail -12
