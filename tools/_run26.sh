set -x
mkdir -p gpurun_out/r2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke26.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2/smoke26.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2/gputest26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest26.log
tail -4 gpurun_out/r2/gputest26.log
timeout 300 python bench.py --skip-secondary --steps 10 > gpurun_out/r2/bench26.json 2> gpurun_out/r2/bench26.err; cut -c1-200 gpurun_out/r2/bench26.json
