mkdir -p gpurun_out/r2s38
O=gpurun_out/r2s38
( time timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1 ) 2> $O/pytest.time; echo "pytest rc=$?" >> $O/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
( time timeout 900 python bench.py > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time; echo "rc=$?" >> $O/bench.err
timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
GCB200_NO_GRAPH=1 timeout 600 python bench.py --profile --steps 2 --warmup 3 > $O/plain.log 2>&1
GCB200_NO_GRAPH=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --profile --steps 2 --warmup 3 > $O/ncu_list.log 2>&1
python tools/launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1
head -c 20000 $O/launches.csv > $O/launches_head.csv; rm -f $O/launches.csv
