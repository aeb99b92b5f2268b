python -m pytest tests/test_rowkernels_gpu.py -x -q -m gpu -k "question_encoder or attention" > gpurun_out/s2l_kern.log 2>&1; tail -4 gpurun_out/s2l_kern.log
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_training_gpu.py tests/test_optim_gpu.py -x -q -m gpu > gpurun_out/s2l_mod.log 2>&1; tail -4 gpurun_out/s2l_mod.log
for rep in 1 2; do
for cfg in "A MCAN_LSTM=1 MCAN_ATTN_BIAS_GRADS=1 MCAN_SPLITK_FWD=1" "B MCAN_LSTM=0 MCAN_ATTN_BIAS_GRADS=1 MCAN_SPLITK_FWD=1" "C MCAN_LSTM=1 MCAN_ATTN_BIAS_GRADS=0 MCAN_SPLITK_FWD=1" "D MCAN_LSTM=1 MCAN_ATTN_BIAS_GRADS=1 MCAN_SPLITK_FWD=0"; do
set -- $cfg; name=$1; shift
env "$@" timeout 300 python bench.py --skip-cpu --steps 30 > gpurun_out/s2l_bench_$name$rep.log 2>&1
echo "$name$rep $* $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/s2l_bench_$name$rep.log | head -1) $(grep -o '"kernels_per_step": [0-9.]*' gpurun_out/s2l_bench_$name$rep.log)"
done; done
