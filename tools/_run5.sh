set -x
mkdir -p gpurun_out/r2
EB200_GEMM_TMA=0 timeout 300 python tools/gemm_bench.py > gpurun_out/r2/gemm5_old.jsonl 2>&1
EB200_GEMM_TMA=1 timeout 300 python tools/gemm_bench.py > gpurun_out/r2/gemm5_tma.jsonl 2>&1
cat gpurun_out/r2/gemm5_old.jsonl gpurun_out/r2/gemm5_tma.jsonl
for tma in 0 1; do for fac in 0 1; do
  EB200_VAR_TMA=$tma EB200_GRID_FACTORISED=$fac timeout 300 python tools/predict_bench.py 4096 9 4096 2 >> gpurun_out/r2/predict5.jsonl 2>> gpurun_out/r2/predict5.err
done; done
cat gpurun_out/r2/predict5.jsonl; tail -3 gpurun_out/r2/predict5.err
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2/gputest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest5.log
tail -15 gpurun_out/r2/gputest5.log
