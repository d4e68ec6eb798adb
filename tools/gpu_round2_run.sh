mkdir -p gpurun_out/r2s24
O=gpurun_out/r2s24
timeout 600 python -m pytest tests/test_anytime_gpu.py tests/test_informed_extend_gpu.py tests/test_sampler_gpu.py tests/test_planner_gpu.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/c5_probe.py 4096 2000 > $O/c5_probe.json 2> $O/c5_probe.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30000 -c 4000 --csv --log-file $O/launches.csv python tools/c5_probe.py 4096 2000 3000 > $O/ncu.log 2>&1
python tools/launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1
