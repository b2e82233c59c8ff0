mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 20 --mode lock > gpurun_out/t33_lock8.json 2> gpurun_out/t33_lock8.err; python -c "
import json;d=json.load(open('gpurun_out/t33_lock8.json'));print('8gpu lock DOUBLE dev',d['value']/1e6,d['ms_per_step']*1e3,'warm',d['warm_cache_ms_per_step']*1e3,'e2e',d['e2e']['value']/1e6,d['e2e']['ms_per_step']*1e3, d.get('master_counters'))"
SPARKFLOW_PUBLISH=single timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 200 --warmup 20 --mode lock > gpurun_out/t33_lock8s.json 2> gpurun_out/t33_lock8s.err; python -c "
import json;d=json.load(open('gpurun_out/t33_lock8s.json'));print('8gpu lock SINGLE dev',d['value']/1e6,d['ms_per_step']*1e3,'warm',d['warm_cache_ms_per_step']*1e3,'e2e',d['e2e']['value']/1e6,d['e2e']['ms_per_step']*1e3)"
