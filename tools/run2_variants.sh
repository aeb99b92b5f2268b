# data-parallel A/B runs (see profiles/*data_parallel_variants.txt); usage: NG=2 bash tools/run2_variants.sh
i=0
run() { i=$((i+1)); name=$1; shift; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port $((29600+i)) bench.py --gpus ${NG:-2} --steps 20 --warmup 5 > gpurun_out/t46_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/t46_$name.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d.get("replica_checksum_divergence"))' 2>&1 | tail -1)"; }
run bucket192 X=1
run perlayer MCAN_DP_BUCKET_MB=0
