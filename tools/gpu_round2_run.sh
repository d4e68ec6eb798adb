mkdir -p gpurun_out/r2s19
O=gpurun_out/r2s19
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_nn_gpu.py -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python tools/nn_bench.py --ops knn --tiles 64,4096 > $O/nn_bench.jsonl 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/nn_launches.csv python tools/nn_bench.py --ops knn --tiles 4096 --reps 1 > $O/ncu_nn.log 2>&1
python tools/launch_summary.py $O/nn_launches.csv > $O/nn_launches_summary.txt 2>&1
