mkdir -p gpurun_out
timeout 300 python tools/trace_e2e.py 2>&1 | tail -3
timeout 300 python bench.py --steps 300 --warmup 30 --mode hogwild > gpurun_out/t19_bench.json 2> gpurun_out/t19_bench.err; python -c "
import json;d=json.load(open('gpurun_out/t19_bench.json'));print('1gpu hog dev',d['value']/1e6,d['ms_per_step']*1e3,'warm',d['warm_cache_ms_per_step']*1e3,'e2e',d['e2e'])"
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" > gpurun_out/t19_gemm_tests.log 2>&1; tail -15 gpurun_out/t19_gemm_tests.log
timeout 300 python tools/bench_gemm.py 2>&1 | tail -8
