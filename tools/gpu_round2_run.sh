mkdir -p gpurun_out/r2s27
O=gpurun_out/r2s27
timeout 900 python -m pytest tests/test_edge_gpu.py tests/test_planner_gpu.py tests/test_golden_gpu.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py --skip-nn --skip-extras > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
GCB200_NO_GRAPH=1 timeout 600 python bench.py --profile --steps 2 --warmup 3 > $O/plain.log 2>&1
GCB200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:arm_grid_kernel -s 60 -c 3 -o $O/armgrid python bench.py --profile --steps 2 --warmup 3 > $O/ncu_full.log 2>&1
python tools/ncu_summary.py $O/armgrid.ncu-rep > $O/armgrid_summary.txt 2>&1
ncu -i $O/armgrid.ncu-rep --page source --csv --print-source cuda,sass > $O/armgrid_source.csv 2>/dev/null
python tools/ncu_lines.py $O/armgrid_source.csv > $O/armgrid_lines.txt 2>&1
gzip -f $O/armgrid_source.csv; rm -f $O/armgrid.ncu-rep
