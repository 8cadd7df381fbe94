set -x
mkdir -p gpurun_out/r2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/bench_n8_final.json 2> gpurun_out/r2/bench_n8_final.err; echo "bench8 rc=$?"
tail -2 gpurun_out/r2/bench_n8_final.err
