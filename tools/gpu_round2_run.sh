mkdir -p gpurun_out/r2s5
O=gpurun_out/r2s5
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
timeout 300 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
timeout 300 python bench.py --profile --steps 2 --warmup 3 > $O/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --profile --steps 2 --warmup 3 > $O/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:arm_grid_kernel -s 400 -c 1 -o $O/armgrid python bench.py --profile --steps 2 --warmup 3 > $O/ncu_full.log 2>&1
ls -la $O
