timeout 300 python -m pytest tests/test_rowkernels_gpu.py tests/test_modules_gpu.py -x -q -m gpu -k "question_encoder or fused_head or golden or full_size" > gpurun_out/s3d_tests.log 2>&1; tail -3 gpurun_out/s3d_tests.log
for rep in 1 2; do
for cfg in "A MCAN_SPLITK_HEAD=1" "B MCAN_SPLITK_HEAD=0"; do
set -- $cfg; name=$1; shift
env "$@" MCAN_BENCH_DUMP=s3d_records_$name$rep.json timeout 300 python bench.py --skip-cpu --steps 30 > gpurun_out/s3d_bench_$name$rep.log 2>&1
echo "$name$rep $* $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/s3d_bench_$name$rep.log | head -1)"
done; done
