mkdir -p gpurun_out/r2s25
O=gpurun_out/r2s25
timeout 600 python -m pytest tests/test_anytime_gpu.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
( time timeout 900 python bench.py > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time; echo "rc=$?" >> $O/bench.err
