set -x
mkdir -p gpurun_out/r2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --sweep-log2-rows 17 > gpurun_out/r2/bench_n8_b.json 2> gpurun_out/r2/bench_n8_b.err; echo "bench8 rc=$?"
tail -3 gpurun_out/r2/bench_n8_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_n8_b.json').read().strip().splitlines()[-1])
l=d["loo_c4"]; print("loo", l["seconds"], l["refits_per_rank"], l["device_evaluations"], l["combined_native_calls"], l["seconds_inside_native_calls_per_rank"])
print("headline", d["value"], "sweep", d["sweep_c5"]["seconds"])
PY
