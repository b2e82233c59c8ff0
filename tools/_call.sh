mkdir -p gpurun_out
tools/run_scaling.sh v3batch 200 20 "1 4 8" > gpurun_out/t27_scaling.log 2>&1; cat gpurun_out/t27_scaling.log
