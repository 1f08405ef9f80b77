#!/usr/bin/env bash
# One gpurun call that refreshes the measured artefacts of a round (run from the repo root ON THE GPU BOX):
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/measure_all.sh r2_b'
#
# Writes everything under gpurun_out/<tag>_*; copy what should be judged into profiles/.
#   1. ncu launch list of a short bench run   (gpu__time_duration per launch: kernel shares of the step)
#   2. ncu --set full of the two pair kernels (fused list build + mass sweep, force replay) + summary tables
#   3. tools/small_configs.py                 (the shipped configs next to the CPU port)
set -u
TAG="${1:-r2}"
OUT=gpurun_out
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/${TAG}_gpu.csv" 2>&1
( nproc; lscpu | grep -E "Model name|^CPU\(s\)" ) > "$OUT/${TAG}_host.txt" 2>&1

SP_BENCH_ALLOW_SHORT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/${TAG}_launches.csv" \
    python bench.py --steps 3 --warmup 1 --no-cpu --no-box > "$OUT/${TAG}_launches_run.log" 2>&1

timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_nbr_build_sweep|k_sweep_list' -s 6 -c 2 \
    -o "$OUT/${TAG}_pair_kernels" -f python tools/sweep_bench.py --steps 1 --warm 5 > "$OUT/${TAG}_ncu_run.log" 2>&1
python tools/ncu_summary.py "$OUT/${TAG}_pair_kernels.ncu-rep" > "$OUT/${TAG}_pair_kernels_ncu.md" 2>&1
python tools/ncu_mem.py "$OUT/${TAG}_pair_kernels.ncu-rep" > "$OUT/${TAG}_pair_kernels_mem.txt" 2>&1

timeout 600 python tools/small_configs.py > "$OUT/${TAG}_small_configs.log" 2>&1
cp -f "$OUT/small_configs.json" "$OUT/${TAG}_small_configs.json" 2>/dev/null
ls -la "$OUT" | tail -20
