set -x
mkdir -p gpurun_out/r2
export EB200_LIB=$PWD/emulator_b200/libemulator_b200_alt.so
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke15.log 2>&1 || { echo "SMOKE FAILED"; tail -20 gpurun_out/r2/smoke15.log; exit 1; }
EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 4096 9 8 5 >> gpurun_out/r2/group15.jsonl 2>> gpurun_out/r2/group15.err
EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 8176 9 4 2 >> gpurun_out/r2/group15.jsonl 2>> gpurun_out/r2/group15.err
EB200_GROUP_SIZE=8 timeout 300 python tools/group_bench.py 2048 9 8 10 >> gpurun_out/r2/group15.jsonl 2>> gpurun_out/r2/group15.err
EB200_GROUP_SIZE=8 timeout 300 python tools/group_bench.py 1000 3 8 20 >> gpurun_out/r2/group15.jsonl 2>> gpurun_out/r2/group15.err
cat gpurun_out/r2/group15.jsonl | cut -c1-330; tail -3 gpurun_out/r2/group15.err
