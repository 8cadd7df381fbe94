set -x
mkdir -p gpurun_out/r2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke17.log 2>&1 || { echo "SMOKE FAILED"; tail -20 gpurun_out/r2/smoke17.log; exit 1; }
for gs in 4 8; do for ch in 1 2 3; do
EB200_GROUP_SIZE=$gs EB200_GROUP_CHAIN=$ch timeout 300 python tools/group_bench.py 4096 9 8 5 >> gpurun_out/r2/group17.jsonl 2>> gpurun_out/r2/group17.err
done; done
EB200_GROUP_SIZE=8 timeout 300 python tools/group_bench.py 2048 9 8 10 >> gpurun_out/r2/group17.jsonl 2>> gpurun_out/r2/group17.err
EB200_GROUP_SIZE=8 timeout 300 python tools/group_bench.py 3072 9 8 5 >> gpurun_out/r2/group17.jsonl 2>> gpurun_out/r2/group17.err
EB200_GROUP_SIZE=2 EB200_GROUP_MAX_NT=200 timeout 300 python tools/group_bench.py 8176 9 4 2 >> gpurun_out/r2/group17.jsonl 2>> gpurun_out/r2/group17.err
EB200_GROUP_SIZE=4 EB200_GROUP_MAX_NT=200 timeout 300 python tools/group_bench.py 8176 9 4 2 >> gpurun_out/r2/group17.jsonl 2>> gpurun_out/r2/group17.err
EB200_GROUP_SIZE=4 timeout 300 python tools/group_bench.py 6144 9 4 3 >> gpurun_out/r2/group17.jsonl 2>> gpurun_out/r2/group17.err
python - <<'PY'
import json
for l in open('gpurun_out/r2/group17.jsonl'):
    d=json.loads(l); print(d["n"], "gs", d["group_size"], "chain", d["chain"], "single %.3f group %.3f bit %s" % (d["single_ms_per_eval"], d["group_ms_per_eval"], d["bit_identical"]))
PY
tail -3 gpurun_out/r2/group17.err
