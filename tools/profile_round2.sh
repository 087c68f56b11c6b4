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
    -k regex:'k_decode_vertex_scan|k_encode_vtx_packed|k_vertex_candidates|k_bounds_reduce|k_requant|k_flatten|k_nocomp|k_scan_prep|k_vertex_order|k_gather_packed|k_wide_collect' \
    -c 40 -o gpurun_out/r02_kernels_10m python tools/profile_driver.py --nr 2237 --ns 4472 --reps 1 > gpurun_out/r02_kernels_10m.log 2>&1
# 3. the batch: launch list and the scan decoder over 195 meshes
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_batch390.csv \
    python tools/batch_probe.py --meshes 390 --reps 1 > gpurun_out/r02_launches_batch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_decode_vertex_scan|k_vertex_candidates_stage|k_flatten|k_global_face_off' -c 4 \
    -o gpurun_out/r02_batch390 python tools/batch_probe.py --meshes 390 --reps 1 > gpurun_out/r02_batch390.log 2>&1
# summaries are made here (same image, same ncu): the reports themselves exceed what gpurun brings back
python tools/summarize_profiles.py launches gpurun_out/r02_launches_bench_10m.csv gpurun_out/r02_launches_bench_10m_summary.csv
python tools/summarize_profiles.py launches gpurun_out/r02_launches_batch390.csv gpurun_out/r02_launches_batch390_summary.csv
python tools/summarize_profiles.py full gpurun_out/r02_kernels_10m.ncu-rep gpurun_out/r02_kernels_10m_ncu_full.txt
python tools/summarize_profiles.py full gpurun_out/r02_batch390.ncu-rep gpurun_out/r02_batch390_ncu_full.txt
for k in k_decode_vertex_scan k_vertex_candidates_stage k_encode_vtx_packed; do
    ncu -i gpurun_out/r02_kernels_10m.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:$k > /tmp/src_$k.csv 2>/dev/null
    python tools/ncu_lines.py /tmp/src_$k.csv 30 > gpurun_out/r02_lines_10m_$k.txt 2>&1
done
ncu -i gpurun_out/r02_batch390.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:k_decode_vertex_scan > /tmp/src_b.csv 2>/dev/null
python tools/ncu_lines.py /tmp/src_b.csv 30 > gpurun_out/r02_lines_batch390_k_decode_vertex_scan.txt 2>&1
rm -f gpurun_out/r02_kernels_10m.ncu-rep gpurun_out/r02_batch390.ncu-rep gpurun_out/r02_launches_bench_10m.csv gpurun_out/r02_launches_batch390.csv
ls -la gpurun_out | tail -14
