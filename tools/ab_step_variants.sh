python -m pytest tests/test_rowkernels_gpu.py -x -q -m gpu -k "attention" > gpurun_out/s3a_kern.log 2>&1; tail -3 gpurun_out/s3a_kern.log
for v in 1 0; do echo "MCAN_ATTN_PREFETCH=$v"; MCAN_ATTN_PREFETCH=$v python tools/attn_bench.py 2>&1 | tail -5; done > gpurun_out/s3a_attn_bench.txt; cat gpurun_out/s3a_attn_bench.txt
for rep in 1 2; do
for cfg in "A MCAN_ATTN_PREFETCH=1" "B MCAN_ATTN_PREFETCH=0"; do
set -- $cfg; name=$1; shift
env "$@" timeout 300 python bench.py --skip-cpu --steps 30 > gpurun_out/s3a_bench_$name$rep.log 2>&1
echo "$name$rep $* $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/s3a_bench_$name$rep.log | head -1)"
done; done
