set -x
mkdir -p gpurun_out/r2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/bench_n8_final.json 2> gpurun_out/r2/bench_n8_final.err; echo "bench8 rc=$?"
cat gpurun_out/r2/bench_n8_final.json | cut -c1-600; tail -3 gpurun_out/r2/bench_n8_final.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2/gputest_multi_final.log 2>&1; tail -5 gpurun_out/r2/gputest_multi_final.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/bench_n2_final.json 2> gpurun_out/r2/bench_n2_final.err; echo "bench2 rc=$?"
cat gpurun_out/r2/bench_n2_final.json | cut -c1-300
