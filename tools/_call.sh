mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -12
SPARKFLOW_PARTITION=resident timeout 300 python bench.py --steps 300 --warmup 30 > gpurun_out/t29_bench_res.json 2> gpurun_out/t29_bench_res.err; tail -3 gpurun_out/t29_bench_res.err; python -c "
import json;d=json.load(open('gpurun_out/t29_bench_res.json'));print('1gpu lock RESIDENT dev',d['value']/1e6,d['ms_per_step']*1e3,'warm',d['warm_cache_ms_per_step']*1e3,'e2e',d['e2e'], d['final_loss'])"
