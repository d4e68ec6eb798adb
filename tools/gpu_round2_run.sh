mkdir -p gpurun_out/r2s15
O=gpurun_out/r2s15
N=${GC_NGPU:-2}
for ex in nccl p2p; do for q in 1 4096; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tools/sharded_knn_bench.py --check --exchange $ex --queries $q >> $O/sharded_${N}gpu_$ex.jsonl 2>> $O/sharded_${N}gpu.err
done; done
tail -c 2000 $O/sharded_${N}gpu.err > $O/err.tail; rm -f $O/sharded_${N}gpu.err
