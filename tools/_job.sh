timeout 200 python -m pytest tests/test_gpu_persistent.py -x -q > gpurun_out/r2_persist_test4.log 2>&1; echo persist rc=$?; tail -3 gpurun_out/r2_persist_test4.log
for ent in 125000 250000 500000 1000000; do
  for p in 1 0; do
    LHVI_PERSISTENT=$p timeout 200 python bench.py --entities $ent --steps 20 --warmup 5 --no-c2f --no-cpu-baseline > gpurun_out/r2_c_p${p}_$ent.json 2> gpurun_out/r2_c_p${p}_$ent.err; echo bench $p $ent rc=$?
  done
done
for ent in 125000 1000000; do timeout 200 python tools/iter_trace.py --entities $ent --iters 3 > gpurun_out/r2_trace_c_$ent.txt 2>&1; done
