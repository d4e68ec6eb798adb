mkdir -p gpurun_out/r2s40
O=gpurun_out/r2s40
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 > $O/bench_2gpu.json 2> $O/bench_2gpu.err ) 2> $O/bench.time; echo "rc=$?" >> $O/bench_2gpu.err
tail -c 1200 $O/bench_2gpu.err > $O/bench_2gpu.err.tail; rm $O/bench_2gpu.err
