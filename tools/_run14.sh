set -x
mkdir -p gpurun_out/r2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke14.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2/smoke14.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2/gputest14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/gputest14.log
tail -6 gpurun_out/r2/gputest14.log
timeout 900 python bench.py > gpurun_out/r2/bench14.json 2> gpurun_out/r2/bench14.err; echo "bench rc=$?"
tail -3 gpurun_out/r2/bench14.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/bench14_ref.json 2> gpurun_out/r2/bench14_ref.err; echo "ref rc=$?"
cat gpurun_out/r2/bench14_ref.json | cut -c1-300
