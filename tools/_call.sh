mkdir -p gpurun_out
timeout 100 python bench.py --steps 200 --warmup 20 2>gpurun_out/t37.err | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('dev',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),'clocks',d['clocks'])"
tail -2 gpurun_out/t37.err | cut -c1-200
