set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2/gputest6a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest6a.log
tail -25 gpurun_out/r2/gputest6a.log
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_emulators.py -m gpu -q > gpurun_out/r2/gputest6b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest6b.log
tail -25 gpurun_out/r2/gputest6b.log
EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 4096 9 8 5 >> gpurun_out/r2/group6.jsonl 2>> gpurun_out/r2/group6.err
EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 8176 9 4 2 >> gpurun_out/r2/group6.jsonl 2>> gpurun_out/r2/group6.err
cat gpurun_out/r2/group6.jsonl; tail -3 gpurun_out/r2/group6.err
timeout 300 python tools/predict_bench.py 4096 9 4096 2 >> gpurun_out/r2/predict6.jsonl 2>> gpurun_out/r2/predict6.err
cat gpurun_out/r2/predict6.jsonl; tail -3 gpurun_out/r2/predict6.err
timeout 600 python bench.py --skip-secondary > gpurun_out/r2/bench6.json 2> gpurun_out/r2/bench6.err; cat gpurun_out/r2/bench6.json | cut -c1-2500; tail -3 gpurun_out/r2/bench6.err
