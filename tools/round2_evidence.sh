python -m pytest tests -x -q -m gpu > gpurun_out/r02_tests.log 2>&1; tail -3 gpurun_out/r02_tests.log
python bench.py > gpurun_out/r02_bench.log 2>&1; tail -c 600 gpurun_out/r02_bench.log; echo
python bench.py --model small --skip-cpu > gpurun_out/r02_bench_small.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_small.log | head -1
python bench.py --model small --ragged prefix --skip-cpu > gpurun_out/r02_bench_small_ragged.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_small_ragged.log | head -1
python tools/profile_step.py large 1 > gpurun_out/r02_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step.py large 1 > gpurun_out/r02_ncu.log 2>&1; tail -1 gpurun_out/r02_ncu.log
python tools/one_round2.py > gpurun_out/r02_one.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|lstm_|attflat_|rowmask|sigmoid' -s 14 -o gpurun_out/r02_round2_kernels python tools/one_round2.py > gpurun_out/r02_ncu2.log 2>&1; tail -2 gpurun_out/r02_ncu2.log; ls -la gpurun_out/*.ncu-rep
python tools/gemm_bench.py > gpurun_out/r02_gemm_bench.log 2>&1; tail -3 gpurun_out/r02_gemm_bench.log
python tools/lstm_bench.py > gpurun_out/r02_lstm_bench.txt 2>&1; cat gpurun_out/r02_lstm_bench.txt
