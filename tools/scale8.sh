i=0
run() { i=$((i+1)); name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29580+i)) bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/s2t_$name.log 2>&1; echo "$name: $(grep -o '"ms_per_step": [0-9.]*\|"gemm_ms_per_step": [0-9.]*\|"kernels_per_step": [0-9]*' gpurun_out/s2t_$name.log | head -3 | tr '\n' ' ')"; }
run dynamic MCAN_GEMM_DYNAMIC=1
run base
run bucket128 MCAN_DP_BUCKET_MB=128
run bucket400 MCAN_DP_BUCKET_MB=400
