#!/usr/bin/env bash
# One gpurun call that refreshes every measured artefact of a round (run from the repo root ON THE GPU BOX):
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/measure_all.sh r2_a'
#
# Writes everything under gpurun_out/<tag>_*; copy what should be judged into profiles/.
#   1. pytest -m gpu                         (parity gate; the rest is meaningless if it is red)
#   1b. tests/pending_gpu_round2.py          (checks written when no GPU time was left)
#   2. bench.py, N = 1                        (value, e2e, roofline, cpu_baseline) and the reference arm
#   3. ncu launch list of a short bench run   (gpu__time_duration per launch: kernel shares of the step)
#   4. ncu --set full of the three pair kernels (k_nbr_build, mass replay, force replay) + summary table
#   5. tools/small_configs.py                 (the nine shipped configs next to the CPU port)
set -u
TAG="${1:-r2}"
OUT=gpurun_out
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/${TAG}_gpu.csv" 2>&1
( nproc; lscpu | grep -E "Model name|^CPU\(s\)" ) > "$OUT/${TAG}_host.txt" 2>&1

timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/${TAG}_pytest_gpu.log" 2>&1
echo "pytest rc=$?" >> "$OUT/${TAG}_pytest_gpu.log"
tail -3 "$OUT/${TAG}_pytest_gpu.log"

# checks written without a GPU (not collected by the suite above): run them, move the green ones into test_*_gpu.py
timeout 600 python -m pytest tests/pending_gpu_round2.py -q -p no:cacheprovider > "$OUT/${TAG}_pytest_pending.log" 2>&1
echo "pending rc=$?" >> "$OUT/${TAG}_pytest_pending.log"
tail -3 "$OUT/${TAG}_pytest_pending.log"

timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/${TAG}_bench_reference.json" 2> "$OUT/${TAG}_bench_reference.err"
timeout 600 python bench.py > "$OUT/${TAG}_bench.json" 2> "$OUT/${TAG}_bench.err"
tail -c 600 "$OUT/${TAG}_bench.json"

timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${TAG}_launches.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu > "$OUT/${TAG}_launches_run.log" 2>&1

timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_nbr_build|k_sweep_list' -s 6 -c 3 \
    -o "$OUT/${TAG}_pair_kernels" -f python tools/sweep_bench.py --steps 1 --warm 5 > "$OUT/${TAG}_ncu_run.log" 2>&1
python tools/ncu_summary.py "$OUT/${TAG}_pair_kernels.ncu-rep" > "$OUT/${TAG}_pair_kernels_ncu.md" 2>&1

timeout 300 python tools/small_configs.py > "$OUT/${TAG}_small_configs.log" 2>&1
cp -f "$OUT/small_configs.json" "$OUT/${TAG}_small_configs.json" 2>/dev/null
ls -la "$OUT" | tail -20
