timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_multi_test2.log 2>&1; echo multi rc=$?; tail -3 gpurun_out/r2_multi_test2.log
P=29520
for N in 8 4 2; do
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo bench N=$N rc=$?
done
P=$((P+1))
LHVI_PERSISTENT=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_graph_n8.json 2> gpurun_out/r2_bench_graph_n8.err; echo bench graph N=8 rc=$?
P=$((P+1))
LHVI_PERSISTENT=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_p1_n2.json 2> gpurun_out/r2_bench_p1_n2.err; echo bench p1 N=2 rc=$?
P=$((P+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n8.json 2> gpurun_out/r2_bench_ref_n8.err; echo ref N=8 rc=$?
