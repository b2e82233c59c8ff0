mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fetch" 2>&1 | tail -5
timeout 400 python -m pytest tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -12
timeout 300 python tools/trace_e2e.py 2>&1 | tail -1 | cut -c1-1500
timeout 300 python bench.py --steps 300 --warmup 30 > gpurun_out/t26_bench.json 2> gpurun_out/t26_bench.err; tail -3 gpurun_out/t26_bench.err; python -c "
import json;d=json.load(open('gpurun_out/t26_bench.json'));print('1gpu lock dev',d['value']/1e6,d['ms_per_step']*1e3,'warm',d['warm_cache_ms_per_step']*1e3,'e2e',d['e2e'], d['final_loss'])"
