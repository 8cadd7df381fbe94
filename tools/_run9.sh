set -x
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2/gputest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest9.log
tail -6 gpurun_out/r2/gputest9.log
timeout 900 python bench.py > gpurun_out/r2/bench9.json 2> gpurun_out/r2/bench9.err; echo "bench rc=$?"
tail -3 gpurun_out/r2/bench9.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_bench9.csv python bench.py --steps 2 --warmup 3 --skip-secondary > gpurun_out/r2/ncu_bench9.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'factor_dataflow|lauum_trace' -c 4 -o /tmp/eval_full9 python tools/ncu_group.py > gpurun_out/r2/ncu_eval9.log 2>&1
ncu -i /tmp/eval_full9.ncu-rep --page raw --csv > gpurun_out/r2/eval_full9_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'var_gemm|predict_grid_mean|grid_axis|var_finish|transpose_upper' -c 10 -o /tmp/predict_full9 python tools/predict_bench.py 4096 9 512 1 > gpurun_out/r2/ncu_predict9.log 2>&1
ncu -i /tmp/predict_full9.ncu-rep --page raw --csv > gpurun_out/r2/predict_full9_raw.csv 2>/dev/null
ls -la gpurun_out/r2 | tail -12; du -sh gpurun_out
