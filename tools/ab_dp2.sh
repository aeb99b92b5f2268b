python -m pytest tests/test_gemm_gpu.py -x -q -m gpu -k grouped > gpurun_out/s2u_gemm.log 2>&1; tail -2 gpurun_out/s2u_gemm.log
python -m pytest tests/test_dp_gpu.py -x -q -m gpu > gpurun_out/s2u_dp.log 2>&1; tail -3 gpurun_out/s2u_dp.log; tail -1 gpurun_out/dp_equivalence.txt
for rep in 1 2; do
for cfg in "direct MCAN_DP_WGRAD_BF16=1" "staged MCAN_DP_WGRAD_BF16=0"; do
set -- $cfg; name=$1; shift
env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/s2u_bench2_$name$rep.log 2>&1
echo "$name$rep $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/s2u_bench2_$name$rep.log | head -1) $(grep -o '"replica_checksum_divergence": [0-9.e-]*\|"kernels_per_step": [0-9]*' gpurun_out/s2u_bench2_$name$rep.log | tr '\n' ' ')"
done; done
