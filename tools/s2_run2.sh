mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/s2_gputest_full2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_gputest_full2.log
tail -4 gpurun_out/s2_gputest_full2.log
timeout 600 python bench.py > gpurun_out/s2_bench2.json 2> gpurun_out/s2_bench2.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/s2_bench2_ref.json 2> gpurun_out/s2_bench2_ref.err; echo "bench ref rc=$?"
timeout 600 python tools/tc_spread_sweep.py > gpurun_out/s2_tc_sweep.txt 2>&1; echo "sweep rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s2_launches2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/s2_ncu_launch2.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loglr_table --launch-skip 4 --launch-count 2 -o gpurun_out/s2_table_full2 python tools/quick_time.py 16384 256 auto > gpurun_out/s2_ncu_full2.log 2>&1; echo "ncu full rc=$?"
CALIB=8 timeout 300 python tools/quick_time.py 16384 256 recurrence,auto > gpurun_out/s2_calib_time.txt 2>&1; echo "calib rc=$?"
