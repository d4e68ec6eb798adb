mkdir -p gpurun_out/r2s18
O=gpurun_out/r2s18
timeout 600 python -m pytest tests/test_sharded_gpu.py -q -m gpu -k "c_abi" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/sharded_capi_check.py > $O/sharded_capi_2gpu_b.jsonl 2> $O/err.log; tail -c 1500 $O/err.log > $O/err.tail; rm -f $O/err.log
