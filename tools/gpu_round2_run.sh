mkdir -p gpurun_out/r2s21
O=gpurun_out/r2s21
timeout 900 python -m pytest tests/test_anytime_gpu.py tests/test_informed_extend_gpu.py tests/test_sampler_gpu.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/c5_probe.py 256 500 > $O/c5_probe.json 2> $O/c5_probe.err
timeout 600 python tools/c5_probe.py 1024 2000 >> $O/c5_probe.json 2>> $O/c5_probe.err
