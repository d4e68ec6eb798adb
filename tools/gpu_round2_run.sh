mkdir -p gpurun_out/r2s41
O=gpurun_out/r2s41
timeout 600 python -m pytest tests/test_planner_gpu.py tests/test_golden_gpu.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py --skip-nn --skip-extras > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
