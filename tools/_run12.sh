set -x
mkdir -p gpurun_out/r2
timeout 200 python tools/group_trace.py 4096 9 4 > gpurun_out/r2/group_trace_4096_4.txt 2>&1
cat gpurun_out/r2/group_trace_4096_4.txt
timeout 200 python tools/group_trace.py 4096 9 1 > gpurun_out/r2/group_trace_4096_1.txt 2>&1
cat gpurun_out/r2/group_trace_4096_1.txt
timeout 200 python tools/group_trace.py 2048 9 8 > gpurun_out/r2/group_trace_2048_8.txt 2>&1
cat gpurun_out/r2/group_trace_2048_8.txt
