timeout 600 python -m pytest tests/test_gpu_compat.py -q > gpurun_out/r2_pytest_compat.log 2>&1; echo compat rc=$?; tail -12 gpurun_out/r2_pytest_compat.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_compat.py > gpurun_out/r2_pytest9.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2_pytest9.log
timeout 200 python bench.py --entities 125000 --steps 20 --warmup 6 --no-configs --no-c2f --no-cpu-baseline > gpurun_out/r2_h_125000.json 2> gpurun_out/r2_h_125000.err; echo bench rc=$?
for p in 1 0; do LHVI_PERSISTENT=$p timeout 200 python bench.py --entities 125000 --dtype float64 --steps 20 --warmup 6 --no-configs --no-c2f --no-cpu-baseline > gpurun_out/r2_h_f64_p$p.json 2> gpurun_out/r2_h_f64_p$p.err; echo bench f64 $p rc=$?; done
