set -x
mkdir -p gpurun_out/r2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke8.log 2>&1 || { echo "SMOKE FAILED"; tail -20 gpurun_out/r2/smoke8.log; exit 1; }
for lib in "" "$PWD/emulator_b200/libemulator_b200_alt.so"; do
  EB200_LIB=$lib EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 4096 9 8 5 >> gpurun_out/r2/group8.jsonl 2>> gpurun_out/r2/group8.err
  EB200_LIB=$lib EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 8176 9 4 2 >> gpurun_out/r2/group8.jsonl 2>> gpurun_out/r2/group8.err
done
cat gpurun_out/r2/group8.jsonl; tail -3 gpurun_out/r2/group8.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2/gputest8a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest8a.log
tail -4 gpurun_out/r2/gputest8a.log
