python -m pytest tests/test_dp_gpu.py -x -q -m gpu > gpurun_out/s2q_dp.log 2>&1; tail -3 gpurun_out/s2q_dp.log; tail -1 gpurun_out/dp_equivalence.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/s2q_bench2.log 2>&1
grep -o '"ms_per_step": [0-9.]*\|"grad_exchange": "[^"]*"' gpurun_out/s2q_bench2.log | head -3
