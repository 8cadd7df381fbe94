set -x
mkdir -p gpurun_out/r2
for fused in 1 0; do EB200_GRID_FUSED=$fused timeout 300 python tools/predict_bench.py 4096 9 4096 2 >> gpurun_out/r2/predict11.jsonl 2>> gpurun_out/r2/predict11.err; done
cat gpurun_out/r2/predict11.jsonl; tail -3 gpurun_out/r2/predict11.err
