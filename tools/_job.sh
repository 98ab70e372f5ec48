timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest8.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2_pytest8.log
timeout 900 python bench.py --steps 20 --warmup 5 --config 2 > gpurun_out/r2_bench_c2f.json 2> gpurun_out/r2_bench_c2f.err; echo bench rc=$?
