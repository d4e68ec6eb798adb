mkdir -p gpurun_out/r2s35
O=gpurun_out/r2s35
timeout 900 python -m pytest tests/test_ref_dropin_gpu.py tests/test_anytime_gpu.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
