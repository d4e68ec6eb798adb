mkdir -p gpurun_out/r2s9
O=gpurun_out/r2s9
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
du -sh $O; ls -la $O
