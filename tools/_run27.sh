set -x
mkdir -p gpurun_out/r2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke27.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2/smoke27.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2/gputest27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest27.log
tail -3 gpurun_out/r2/gputest27.log
timeout 300 python bench.py --skip-secondary --steps 10 > gpurun_out/r2/bench27.json 2> gpurun_out/r2/bench27.err; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench27.json').read().strip().splitlines()[-1]); print('headline', d['value'], d['roofline']['stage_ms_per_eval_grouped'])"
timeout 200 python tools/predict_bench.py 4096 9 4096 2
