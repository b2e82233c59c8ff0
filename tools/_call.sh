mkdir -p gpurun_out
SPARKFLOW_MEGAKERNEL=1 timeout 60 python bench.py --steps 200 --warmup 20 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('MEGA dev',round(d['ms_per_step']*1e3,1),'warm',round(d['warm_cache_ms_per_step']*1e3,1),'e2e',round(d['e2e']['ms_per_step']*1e3,1), d['config']['kernels_per_step'], d['final_loss'])"
timeout 60 python bench.py --steps 200 --warmup 20 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('BASE dev',round(d['ms_per_step']*1e3,1),'warm',round(d['warm_cache_ms_per_step']*1e3,1),'e2e',round(d['e2e']['ms_per_step']*1e3,1))"
