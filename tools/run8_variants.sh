# data-parallel A/B runs at 8 GPUs (see profiles/*data_parallel_variants.txt); usage: bash tools/run8_variants.sh
i=0
run() { i=$((i+1)); name=$1; shift; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29580+i)) bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/t43_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/t43_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d.get("replica_checksum_divergence"))' 2>&1 | tail -1)"; }
run nopdl_static MCAN_PDL=0
run nopdl_dynamic MCAN_PDL=0 MCAN_GEMM_DYNAMIC=1
