# Round-end GPU record (one B200): the -m gpu tests, smoke(), the default bench line and the reference arm, then the ncu
# launch list of a two-commit run at bench.py's table budget (DRAM bytes per launch -> scripts/update_traffic.py) and one full
# capture of the round kernels.  Every profiler pass runs only after the same program has exited 0 without ncu.
set -x
TAG=${1:-r2f}
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -4 gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 400 gpurun_out/${TAG}_bench_ref.json
python scripts/commit_dev_once.py 1024 1024 2 70000 > gpurun_out/${TAG}_once.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/commit_dev_once.py 1024 1024 2 70000 > gpurun_out/${TAG}_ncu_once.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bat_finish|k_bat_prefix" --launch-skip 20 -c 6 -o gpurun_out/${TAG}_bat -f python scripts/commit_dev_once.py 1024 1024 2 70000 > gpurun_out/${TAG}_ncu_bat.log 2>&1
ncu -i gpurun_out/${TAG}_bat.ncu-rep --page raw --csv > gpurun_out/${TAG}_bat_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
