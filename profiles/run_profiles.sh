#!/usr/bin/env bash
# Regenerates the profiles/ evidence of one round in ONE gpurun call (B200_PROFILING.md recipe):
#
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/run_profiles.sh r02'
#   cp gpurun_out/r02_* profiles/ && python profiles/make_traffic.py cfg2=profiles/r02_ncu_full_cfg2.csv T=profiles/r02_ncu_full_T.csv
#
# For each workload (cfg2 = bench default, T = north-star target): (1) the bench itself, no profiler - the only source of
# bench numbers; (2) the ncu launch list of the same command (cold-cache, serialised: shares, not absolutes);
# (3) one `--set full` capture of one step's kernels, dumped to CSV with the DRAM byte counters bench.py's roofline cites.
# (4) the accumulation-chain probe (gradient error of a full-size step vs the float64 oracle).
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p "$OUT"
for WL in cfg2 T; do
    python bench.py --workload "$WL" --steps 100 --warmup 5 > "$OUT/${TAG}_bench_${WL}.json" 2> "$OUT/${TAG}_bench_${WL}.err" || { echo "bench $WL failed"; tail -5 "$OUT/${TAG}_bench_${WL}.err"; continue; }
    tail -c 600 "$OUT/${TAG}_bench_${WL}.json"; echo
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${TAG}_launches_${WL}.csv" \
        python bench.py --workload "$WL" --steps 6 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_ncu_${WL}.log" 2>&1
    python profiles/summarize_launches.py "$OUT/${TAG}_launches_${WL}.csv" 40 > "$OUT/${TAG}_launches_${WL}_summary.txt" 2>&1
    # one step's worth of kernels after the bind-time sorts and the warm-up steps (launch-skip tuned to the launch list)
    ncu --set full --clock-control none --import-source on --launch-skip 200 -c 40 -o "$OUT/${TAG}_full_${WL}" -f \
        python bench.py --workload "$WL" --steps 6 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_ncu_full_${WL}.log" 2>&1
    ncu -i "$OUT/${TAG}_full_${WL}.ncu-rep" --page raw --csv \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > "$OUT/${TAG}_ncu_full_${WL}.csv" 2>/dev/null
done
python tests/probe_accum_chain.py T default > "$OUT/${TAG}_accum_chain.jsonl" 2>&1
ls -la "$OUT" | tail -20
