mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 --impl nccl > gpurun_out/t31_nccl2.json 2> gpurun_out/t31_nccl2.err; tail -1 gpurun_out/t31_nccl2.json | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/sweep_pushpull.py --max-mb 256 > gpurun_out/t31_sweep.log 2>&1; tail -12 gpurun_out/t31_sweep.log | cut -c1-400
timeout 200 python bench.py --model wide_dnn --steps 30 --warmup 5 --mode hogwild 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('wide_dnn dev %.1fus e2e %.1fus' % (d['ms_per_step']*1e3,d['e2e']['ms_per_step']*1e3))"
