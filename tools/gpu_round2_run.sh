# round-2 evidence run (see profiles/r02_*): launch list and ncu capture of the lock-step step (the bench brackets its
# timed steps with cudaProfilerStart/Stop: --profile-from-start off; the CUDA graph is disabled because ncu mislabels
# graph kernel nodes).  .ncu-rep files are digested on the box and removed: gpurun_out/ travels back only under 64 MiB.
mkdir -p gpurun_out/r2s7
O=gpurun_out/r2s7
digest() {  # $1 = report stem
  python tools/ncu_summary.py $O/$1.ncu-rep > $O/$1_summary.txt 2>&1
  ncu -i $O/$1.ncu-rep --page source --csv --print-source cuda,sass > $O/$1_source.csv 2>/dev/null
  python tools/ncu_lines.py $O/$1_source.csv > $O/$1_lines.txt 2>&1
  gzip -f $O/$1_source.csv
  rm -f $O/$1.ncu-rep
}
timeout 1200 python -m pytest tests/test_birrt_gpu.py tests/test_ref_dropin_gpu.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
export GCB200_NO_GRAPH=1
timeout 300 python bench.py --profile --steps 2 --warmup 3 > $O/plain.log 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv python bench.py --profile --steps 2 --warmup 3 > $O/ncu_list.log 2>&1
python tools/launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1
head -n 400 $O/launches.csv > $O/launches_head.csv; rm -f $O/launches.csv
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:arm_grid_kernel -s 60 -c 3 -o $O/armgrid python bench.py --profile --steps 2 --warmup 3 > $O/ncu_full.log 2>&1
digest armgrid
unset GCB200_NO_GRAPH
du -sh $O; ls -la $O
