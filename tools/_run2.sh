set -x
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gputest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest2.log
tail -30 gpurun_out/r2/gputest2.log
timeout 900 python bench.py > gpurun_out/r2/bench2.json 2> gpurun_out/r2/bench2.err; echo "bench rc=$?"
cat gpurun_out/r2/bench2.json; tail -5 gpurun_out/r2/bench2.err
for ch in 4096 8192 16384 32768; do
  EB200_PREDICT_CHUNK=$ch timeout 300 python tools/predict_bench.py 4096 9 4096 2 >> gpurun_out/r2/predict2.jsonl 2>> gpurun_out/r2/predict2.err
done
for grp in 4 8 16; do
  EB200_PREDICT_CHUNK=16384 EB200_VAR_GROUP=$grp timeout 300 python tools/predict_bench.py 4096 9 4096 2 >> gpurun_out/r2/predict2.jsonl 2>> gpurun_out/r2/predict2.err
done
EB200_PREDICT_CHUNK=32768 EB200_VAR_GROUP=8 timeout 300 python tools/predict_bench.py 4096 9 4096 2 >> gpurun_out/r2/predict2.jsonl 2>> gpurun_out/r2/predict2.err
cat gpurun_out/r2/predict2.jsonl; tail -3 gpurun_out/r2/predict2.err
# ncu: launch list of the headline, then full captures of the dominant kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-secondary > gpurun_out/r2/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'factor_dataflow|lauum_trace' -c 4 -o gpurun_out/r2/eval_full python tools/group_bench.py 4096 9 4 1 > gpurun_out/r2/ncu_eval.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'var_gemm|predict_mean|grid_rows|var_finish' -c 8 -o gpurun_out/r2/predict_full python tools/predict_bench.py 4096 9 512 1 > gpurun_out/r2/ncu_predict.log 2>&1
ls -la gpurun_out/r2
