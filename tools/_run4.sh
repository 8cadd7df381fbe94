set -x
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/bench_n8.json 2> gpurun_out/r2/bench_n8.err; echo "bench8 rc=$?"
cat gpurun_out/r2/bench_n8.json; tail -5 gpurun_out/r2/bench_n8.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/c4_full.py 512 > gpurun_out/r2/c4_full_n8.json 2> gpurun_out/r2/c4_full_n8.err; echo "c4 rc=$?"
cat gpurun_out/r2/c4_full_n8.json; tail -5 gpurun_out/r2/c4_full_n8.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2/gputest_multi.log 2>&1; tail -3 gpurun_out/r2/gputest_multi.log
