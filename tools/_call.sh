mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t34_all.log 2>&1; tail -5 gpurun_out/t34_all.log
timeout 200 python -c "
import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/t34_bench_default.json 2> gpurun_out/t34_bench_default.err; python -c "
import json;d=json.load(open('gpurun_out/t34_bench_default.json'));print('default bench dev',d['value']/1e6,d['ms_per_step']*1e3,'e2e',d['e2e']['value']/1e6,d['e2e']['ms_per_step']*1e3, d['gpu_launches'], d['clocks'])"
timeout 60 python bench.py --impl reference
