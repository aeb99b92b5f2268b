rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -x -q -m gpu > gpurun_out/r02b_tests.log 2>&1; tail -3 gpurun_out/r02b_tests.log
grep top1 gpurun_out/parity_report.jsonl | cut -c1-330
for cfg in "split MCAN_LSTM_SPLIT_INPUT=1" "plain MCAN_LSTM_SPLIT_INPUT=0"; do
set -- $cfg; name=$1; shift
env "$@" timeout 300 python bench.py --skip-cpu --steps 30 > gpurun_out/r02b_bench_$name.log 2>&1
echo "$name $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02b_bench_$name.log | head -1)"
done
timeout 300 python bench.py --skip-cpu --steps 30 --no-graph > gpurun_out/r02b_bench_eager.log 2>&1; echo "eager $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02b_bench_eager.log | head -1)"
python __graft_entry__.py smoke 2>&1 | tail -1
