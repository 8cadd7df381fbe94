set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gputest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest1.log
tail -5 gpurun_out/r2/gputest1.log
for gs in 2 4 8; do
  EB200_GROUP_SIZE=$gs timeout 300 python tools/group_bench.py 4096 9 8 5 >> gpurun_out/r2/group1.jsonl 2>> gpurun_out/r2/group1.err
done
for ch in 1 2 3 6; do
  EB200_GROUP_SIZE=8 EB200_GROUP_CHAIN=$ch timeout 300 python tools/group_bench.py 4096 9 8 5 >> gpurun_out/r2/group1.jsonl 2>> gpurun_out/r2/group1.err
  EB200_GROUP_SIZE=4 EB200_GROUP_CHAIN=$ch timeout 300 python tools/group_bench.py 4096 9 8 5 >> gpurun_out/r2/group1.jsonl 2>> gpurun_out/r2/group1.err
done
EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 8176 9 4 2 >> gpurun_out/r2/group1.jsonl 2>> gpurun_out/r2/group1.err
EB200_GROUP_SIZE=2 timeout 300 python tools/group_bench.py 8176 9 4 2 >> gpurun_out/r2/group1.jsonl 2>> gpurun_out/r2/group1.err
EB200_GROUP_SIZE=8 timeout 300 python tools/group_bench.py 2048 9 8 10 >> gpurun_out/r2/group1.jsonl 2>> gpurun_out/r2/group1.err
cat gpurun_out/r2/group1.jsonl
tail -5 gpurun_out/r2/group1.err
