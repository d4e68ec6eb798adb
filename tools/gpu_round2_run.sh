mkdir -p gpurun_out/r2s17
O=gpurun_out/r2s17
timeout 900 python -m pytest tests/test_sharded_gpu.py -q -m gpu -k "c_abi" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
