rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -x -q -m gpu > gpurun_out/r02c_tests.log 2>&1; tail -3 gpurun_out/r02c_tests.log
MCAN_BENCH_DUMP=r02c_gemm_records.json timeout 300 python bench.py --skip-cpu --steps 30 > gpurun_out/r02c_bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*\|"achieved": [0-9.]*' gpurun_out/r02c_bench.log | head -3
python __graft_entry__.py smoke 2>&1 | tail -1
