# development aid: the full single-GPU measurement pass behind profiles/r2_* (GPU tests, bench, reference arm, tc sweep,
# ncu launch list + full capture of one step's loglr launches, calibrated timing, C4, PT timeline)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/s3_gputest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_gputest_full.log
tail -4 gpurun_out/s3_gputest_full.log
timeout 600 python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/s3_bench_ref.json 2> gpurun_out/s3_bench_ref.err; echo "bench ref rc=$?"
timeout 600 python tools/tc_spread_sweep.py > gpurun_out/s3_tc_sweep.txt 2>&1; echo "sweep rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/s3_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:loglr_(table|recur)" --launch-skip 9 --launch-count 3 -o gpurun_out/s3_step_full python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/s3_ncu_full.log 2>&1; echo "ncu full rc=$?"
CALIB=8 timeout 300 python tools/quick_time.py 16384 256 recurrence,auto > gpurun_out/s3_calib_time.txt 2>&1; echo "calib rc=$?"
timeout 300 python tools/run_pt_c4.py 8 2048 400 256 device 2>&1 | grep "wall\|C4 PT" > gpurun_out/s3_c4_n1.txt; echo "c4 rc=$?"
timeout 300 python tools/profile_pt.py 8 2>&1 | grep -v "Warn\|warn" | tail -32 > gpurun_out/s3_profile_pt_n1.txt; echo "prof rc=$?"
