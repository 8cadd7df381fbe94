set -x
mkdir -p gpurun_out/r2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke18.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2/smoke18.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2/gputest18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest18.log
tail -4 gpurun_out/r2/gputest18.log
timeout 900 python bench.py > gpurun_out/r2/bench18.json 2> gpurun_out/r2/bench18.err; echo "bench rc=$?"
tail -2 gpurun_out/r2/bench18.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_bench18.csv python bench.py --steps 2 --warmup 3 --skip-secondary > gpurun_out/r2/ncu_bench18.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'factor_dataflow|lauum_trace' -c 4 -o /tmp/eval_full18 python tools/ncu_group.py > gpurun_out/r2/ncu_eval18.log 2>&1
ncu -i /tmp/eval_full18.ncu-rep --page raw --csv > gpurun_out/r2/eval_full18_raw.csv 2>/dev/null
timeout 200 python tools/group_trace.py 4096 9 8 > gpurun_out/r2/group_trace18_4096_8.txt 2>&1
cat gpurun_out/r2/group_trace18_4096_8.txt | head -12
