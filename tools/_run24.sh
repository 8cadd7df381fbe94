set -x
mkdir -p gpurun_out/r2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/c4_full.py 512 > gpurun_out/r2/c4_full_n8_final.json 2> gpurun_out/r2/c4_full_n8_final.err; echo "c4 rc=$?"
cat gpurun_out/r2/c4_full_n8_final.json; tail -3 gpurun_out/r2/c4_full_n8_final.err
