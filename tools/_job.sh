timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest3.log 2>&1; echo pytest rc=$?; tail -12 gpurun_out/r2_pytest3.log
for ent in 125000 250000; do
  timeout 200 python bench.py --entities $ent --steps 20 --warmup 6 --no-c2f --no-cpu-baseline > gpurun_out/r2_d_$ent.json 2> gpurun_out/r2_d_$ent.err; echo bench $ent rc=$?
done
timeout 200 python tools/iter_trace.py --entities 125000 --iters 3 > gpurun_out/r2_trace_d_125000.txt 2>&1
