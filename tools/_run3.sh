set -x
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2/gputest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest3.log
tail -40 gpurun_out/r2/gputest3.log
timeout 900 python bench.py --loo-refits 8 > gpurun_out/r2/bench3.json 2> gpurun_out/r2/bench3.err; echo "bench rc=$?"
cat gpurun_out/r2/bench3.json; tail -5 gpurun_out/r2/bench3.err
for ch in 4096 16384 32768 65536; do
  EB200_PREDICT_CHUNK=$ch timeout 300 python tools/predict_bench.py 4096 9 4096 2 >> gpurun_out/r2/predict3.jsonl 2>> gpurun_out/r2/predict3.err
done
EB200_PREDICT_CHUNK=16384 EB200_VAR_GROUP=8 timeout 300 python tools/predict_bench.py 4096 9 4096 2 >> gpurun_out/r2/predict3.jsonl 2>> gpurun_out/r2/predict3.err
EB200_PREDICT_CHUNK=32768 EB200_VAR_GROUP=8 timeout 300 python tools/predict_bench.py 4096 9 4096 2 >> gpurun_out/r2/predict3.jsonl 2>> gpurun_out/r2/predict3.err
cat gpurun_out/r2/predict3.jsonl; tail -3 gpurun_out/r2/predict3.err
BENCH_STREAMS=2 EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 4096 9 8 5 >> gpurun_out/r2/group3.jsonl 2>> gpurun_out/r2/group3.err
BENCH_STREAMS=2 EB200_GROUP_SIZE=2 timeout 300 python tools/group_bench.py 4096 9 8 5 >> gpurun_out/r2/group3.jsonl 2>> gpurun_out/r2/group3.err
EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 6144 9 4 3 >> gpurun_out/r2/group3.jsonl 2>> gpurun_out/r2/group3.err
EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 3072 9 8 5 >> gpurun_out/r2/group3.jsonl 2>> gpurun_out/r2/group3.err
EB200_GROUP_SIZE=8 timeout 300 python tools/group_bench.py 3072 9 8 5 >> gpurun_out/r2/group3.jsonl 2>> gpurun_out/r2/group3.err
cat gpurun_out/r2/group3.jsonl; tail -3 gpurun_out/r2/group3.err
# ncu: launch list of the headline, then full captures of the dominant kernels (reports stay in /tmp; CSV pages come back)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-secondary > gpurun_out/r2/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'factor_dataflow|lauum_trace' -c 4 -o /tmp/eval_full python tools/group_bench.py 4096 9 4 1 > gpurun_out/r2/ncu_eval.log 2>&1
ncu -i /tmp/eval_full.ncu-rep --page raw --csv > gpurun_out/r2/eval_full_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'var_gemm|predict_mean|grid_rows|var_finish' -c 8 -o /tmp/predict_full python tools/predict_bench.py 4096 9 512 1 > gpurun_out/r2/ncu_predict.log 2>&1
ncu -i /tmp/predict_full.ncu-rep --page raw --csv > gpurun_out/r2/predict_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2 /tmp/*.ncu-rep
du -sh gpurun_out
