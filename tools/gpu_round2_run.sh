mkdir -p gpurun_out/r2s30
O=gpurun_out/r2s30
timeout 300 python -m pytest tests/test_anytime_gpu.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for k in 2 2 4 4; do timeout 300 python tools/c5_probe.py 4096 2000 0 $k >> $O/c5_probe.json 2>> $O/c5_probe.err; done
