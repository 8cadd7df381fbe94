set -x
mkdir -p gpurun_out/r2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke7.log 2>&1 || { echo "SMOKE FAILED"; tail -20 gpurun_out/r2/smoke7.log; exit 1; }
tail -2 gpurun_out/r2/smoke7.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2/gputest7a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest7a.log
tail -25 gpurun_out/r2/gputest7a.log
grep -q "rc=0" gpurun_out/r2/gputest7a.log || exit 1
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_emulators.py -m gpu -q > gpurun_out/r2/gputest7b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest7b.log
tail -8 gpurun_out/r2/gputest7b.log
EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 4096 9 8 5 >> gpurun_out/r2/group7.jsonl 2>> gpurun_out/r2/group7.err
EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 8176 9 4 2 >> gpurun_out/r2/group7.jsonl 2>> gpurun_out/r2/group7.err
EB200_GROUP_SIZE=8 timeout 300 python tools/group_bench.py 1000 3 8 20 >> gpurun_out/r2/group7.jsonl 2>> gpurun_out/r2/group7.err
cat gpurun_out/r2/group7.jsonl; tail -3 gpurun_out/r2/group7.err
timeout 600 python bench.py --skip-secondary > gpurun_out/r2/bench7.json 2> gpurun_out/r2/bench7.err; cat gpurun_out/r2/bench7.json | cut -c1-400; tail -3 gpurun_out/r2/bench7.err
for fac in 0 1; do EB200_GRID_FACTORISED=$fac timeout 300 python tools/predict_bench.py 4096 9 4096 2 >> gpurun_out/r2/predict7.jsonl 2>> gpurun_out/r2/predict7.err; done
cat gpurun_out/r2/predict7.jsonl; tail -3 gpurun_out/r2/predict7.err
