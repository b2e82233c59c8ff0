mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/t17_tests.log 2>&1; tail -4 gpurun_out/t17_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sf_gemm_kernel -s 4 -c 1 -f -o gpurun_out/prof_gemm8192 python tools/profile_gemm.py 8192 8192 8192 > gpurun_out/t17_ncu_gemm.log 2>&1; tail -2 gpurun_out/t17_ncu_gemm.log
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:gemm|push|pull|cast' -s 44 -c 11 -f -o gpurun_out/prof_step python tools/profile_step.py 6 > gpurun_out/t17_ncu_step.log 2>&1; tail -2 gpurun_out/t17_ncu_step.log
timeout 300 python bench.py --steps 300 --warmup 30 > gpurun_out/t17_bench_lock.json 2> gpurun_out/t17_bench_lock.err; cut -c1-400 gpurun_out/t17_bench_lock.json
timeout 300 python bench.py --steps 300 --warmup 30 --mode hogwild > gpurun_out/t17_bench_hog.json 2> gpurun_out/t17_bench_hog.err; cut -c1-400 gpurun_out/t17_bench_hog.json
ls -la gpurun_out/*.ncu-rep
