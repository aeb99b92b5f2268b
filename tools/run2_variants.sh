i=0
run() { i=$((i+1)); name=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port $((29540+i)) bench.py --gpus ${NG:-2} --steps 20 --warmup 5 > gpurun_out/t34_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/t34_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"])' 2>&1 | tail -1)"; }
run base X=1
run maxctas4 NCCL_MAX_CTAS=4
run delay MCAN_DP_DELAY_DECODER=1
run delay_ctas8 MCAN_DP_DELAY_DECODER=1 NCCL_MAX_CTAS=8
run dynamic MCAN_GEMM_DYNAMIC=1
