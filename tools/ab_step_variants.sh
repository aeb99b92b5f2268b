timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu > gpurun_out/s3h_tests.log 2>&1; tail -2 gpurun_out/s3h_tests.log
for rep in 1 2; do
for cfg in "A MCAN_GEMM_SHORT_M_RULE=1" "B MCAN_GEMM_SHORT_M_RULE=0"; do
set -- $cfg; name=$1; shift
env "$@" timeout 300 python bench.py --skip-cpu --steps 30 > gpurun_out/s3h_bench_$name$rep.log 2>&1
echo "$name$rep $* $(grep -o '"ms_per_step": [0-9.]*\|"achieved": [0-9.]*' gpurun_out/s3h_bench_$name$rep.log | head -3 | tr '\n' ' ')"
done; done
