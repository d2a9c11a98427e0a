set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_s3.log 2>&1; tail -3 gpurun_out/pytest_gpu_s3.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s3.log 2>&1; tail -2 gpurun_out/smoke_s3.log
python bench.py > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err; tail -c 600 gpurun_out/bench_s3.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_s3_ref.json 2> gpurun_out/bench_s3_ref.err; tail -c 400 gpurun_out/bench_s3_ref.json
python scripts/bench_small_commit.py > gpurun_out/small_commit.log 2>&1; tail -1 gpurun_out/small_commit.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s3_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prove > gpurun_out/ncu_bench_s3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ba_finish|k_ba_prefix" --launch-skip 72 -c 4 -o gpurun_out/ba_s3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prove > gpurun_out/ncu_ba_s3.log 2>&1
ncu -i gpurun_out/ba_s3.ncu-rep --page raw --csv > gpurun_out/ba_s3_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
