# Final evidence of round 2 after the tcgen05 attention kernels (one gpurun call):
#   bash tools/round2b_evidence.sh
rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -x -q -m gpu > gpurun_out/r02d_tests.log 2>&1; tail -3 gpurun_out/r02d_tests.log
MCAN_BENCH_DUMP=r02d_gemm_records.json timeout 300 python bench.py --skip-cpu --steps 30 > gpurun_out/r02d_bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*\|"achieved": [0-9.]*' gpurun_out/r02d_bench.log | head -3
timeout 200 python bench.py --model small --skip-cpu --steps 30 > gpurun_out/r02d_bench_small.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02d_bench_small.log | head -1
python tools/profile_step.py large 1 > gpurun_out/r02d_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02d_launches.csv python tools/profile_step.py large 1 > gpurun_out/r02d_ncu.log 2>&1; tail -1 gpurun_out/r02d_ncu.log
python tools/one_attn.py > gpurun_out/r02d_one_attn.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -o gpurun_out/r02d_attention_tc python tools/one_attn.py > gpurun_out/r02d_ncu2.log 2>&1; tail -2 gpurun_out/r02d_ncu2.log
python tools/attn_bench.py > gpurun_out/r02d_attn_bench.log 2>&1; cat gpurun_out/r02d_attn_bench.log
python __graft_entry__.py smoke 2>&1 | tail -1
