#!/usr/bin/env bash
# Regenerates the profiles/ evidence of one round in ONE gpurun call (B200_PROFILING.md recipe):
#
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash profiles/run_profiles.sh r02'
#   cp gpurun_out/r02_* profiles/ && python profiles/make_traffic.py r02 T=profiles/r02_ncu_full_T.csv
#
# (1) the bench itself (default workload T, cfg2 rides along as extra.cfg2), no profiler - the only source of bench
#     numbers; (2) the ncu launch lists of T and cfg2 (cold-cache, serialised: shares, not absolutes); (3) one `--set full`
#     capture of one steady-state step's main kernels at T, dumped to CSV (DRAM byte counters, tensor-pipe activity,
#     occupancy) - the .ncu-rep itself is dropped when it exceeds the 64 MiB return limit.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p "$OUT"
python bench.py --steps 50 --warmup 5 > "$OUT/${TAG}_bench_T.json" 2> "$OUT/${TAG}_bench_T.err" || { echo "bench failed"; tail -5 "$OUT/${TAG}_bench_T.err"; }
tail -c 400 "$OUT/${TAG}_bench_T.json"; echo
for WL in T cfg2; do
    ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file "$OUT/${TAG}_launches_${WL}.csv" \
        python bench.py --workload "$WL" --steps 6 --warmup 3 --no-cpu-baseline --no-check --no-extra --min-time 0 > "$OUT/${TAG}_ncu_${WL}.log" 2>&1
    python profiles/summarize_launches.py "$OUT/${TAG}_launches_${WL}.csv" 40 > "$OUT/${TAG}_launches_${WL}_summary.txt" 2>&1
done
# one steady-state step's worth of the main kernels (the filter skips the bind-time sorts: ~115 k_radix_sort<0> launches)
ncu --set full --clock-control none -k "regex:k_tc_bilinear|k_tc_dq|k_tc_dc|k_rows_chunk|k_entity_long2|k_w_long2|k_score|k_encoder_forward_v4|k_dense_finalize|k_radix_sort|k_tc_combine|k_tc_transpose|k_tc_bwd_finish|k_tc_prep" \
    --launch-skip 330 -c 20 -o "$OUT/${TAG}_full_T" -f \
    python bench.py --workload T --steps 6 --warmup 3 --no-cpu-baseline --no-check --no-extra --min-time 0 > "$OUT/${TAG}_ncu_full_T.log" 2>&1
ncu -i "$OUT/${TAG}_full_T.ncu-rep" --page raw --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum \
    > "$OUT/${TAG}_ncu_full_T.csv" 2>/dev/null
S=$(stat -c %s "$OUT/${TAG}_full_T.ncu-rep" 2>/dev/null || echo 0); [ "$S" -gt 30000000 ] && rm -f "$OUT/${TAG}_full_T.ncu-rep"
ls -la "$OUT" | tail -12
