mkdir -p gpurun_out/r2s42
O=gpurun_out/r2s42
( time timeout 600 python bench.py > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time; echo "rc=$?" >> $O/bench.err
