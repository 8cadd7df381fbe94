set -x
mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/c4_full.py 128 > gpurun_out/r2/c4_128_n8.json 2> gpurun_out/r2/c4_128_n8.err; echo "c4 rc=$?"
cat gpurun_out/r2/c4_128_n8.json | cut -c1-900; tail -3 gpurun_out/r2/c4_128_n8.err
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
