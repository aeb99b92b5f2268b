python -m pytest tests/test_dp_gpu.py -x -q -m gpu > gpurun_out/s2p_dp.log 2>&1; tail -3 gpurun_out/s2p_dp.log; cat gpurun_out/dp_equivalence.txt | tail -3
for rep in 1 2; do
for cfg in "fp32 MCAN_DP_COMPRESS=" "bf16 MCAN_DP_COMPRESS=bf16"; do
set -- $cfg; name=$1; shift
env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/s2p_bench2_$name$rep.log 2>&1
echo "$name$rep $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/s2p_bench2_$name$rep.log | head -1) $(grep -o '"replica_checksum_divergence": [0-9.e-]*' gpurun_out/s2p_bench2_$name$rep.log)"
done; done
timeout 300 python bench.py --skip-cpu --steps 30 > gpurun_out/s2p_bench1.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/s2p_bench1.log | head -1
