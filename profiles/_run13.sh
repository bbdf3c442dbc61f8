cd $GRAFT_REPO_ROOT
CMD="python bench.py --workload T --steps 6 --warmup 3 --no-cpu-baseline --no-check --no-extra --min-time 0"
$CMD > gpurun_out/t13_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/t13_launches_T.csv $CMD > gpurun_out/t13_ncu.log 2>&1
echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/t13_launches_T.csv 45
