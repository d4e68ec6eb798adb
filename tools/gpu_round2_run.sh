mkdir -p gpurun_out/r2s34
O=gpurun_out/r2s34
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 > $O/bench_8gpu.json 2> $O/bench_8gpu.err ) 2> $O/bench.time; echo "rc=$?" >> $O/bench_8gpu.err
tail -c 1500 $O/bench_8gpu.err > $O/bench_8gpu.err.tail; rm $O/bench_8gpu.err
