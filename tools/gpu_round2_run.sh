mkdir -p gpurun_out/r2s43
O=gpurun_out/r2s43
timeout 150 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "pytest rc=$?" >> $O/pytest_all.log
