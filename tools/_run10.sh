set -x
mkdir -p gpurun_out/r2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke10.log 2>&1 || { echo "SMOKE FAILED"; tail -20 gpurun_out/r2/smoke10.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "grid or predict or sobol or prediction" > gpurun_out/r2/gputest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest10.log
tail -12 gpurun_out/r2/gputest10.log
for fused in 0 1; do EB200_GRID_FUSED=$fused timeout 300 python tools/predict_bench.py 4096 9 4096 2 >> gpurun_out/r2/predict10.jsonl 2>> gpurun_out/r2/predict10.err; done
EB200_GRID_FUSED=1 timeout 300 python tools/predict_bench.py 8176 9 2048 1 >> gpurun_out/r2/predict10.jsonl 2>> gpurun_out/r2/predict10.err
EB200_GRID_FUSED=0 timeout 300 python tools/predict_bench.py 8176 9 2048 1 >> gpurun_out/r2/predict10.jsonl 2>> gpurun_out/r2/predict10.err
cat gpurun_out/r2/predict10.jsonl; tail -3 gpurun_out/r2/predict10.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'var_grid_fused|predict_grid_mean' -c 4 -o /tmp/fused10 python tools/predict_bench.py 4096 9 512 1 > gpurun_out/r2/ncu_fused10.log 2>&1
ncu -i /tmp/fused10.ncu-rep --page raw --csv > gpurun_out/r2/fused10_raw.csv 2>/dev/null
ls -la gpurun_out/r2/fused10_raw.csv
