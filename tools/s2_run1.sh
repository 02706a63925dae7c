mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_gputest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_gputest_full.log
tail -5 gpurun_out/s2_gputest_full.log
timeout 600 python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/s2_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loglr_table --launch-skip 4 --launch-count 2 -o gpurun_out/s2_table_full python tools/quick_time.py 16384 256 auto > gpurun_out/s2_ncu_full.log 2>&1; echo "ncu full rc=$?"
