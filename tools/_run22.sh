set -x
mkdir -p gpurun_out/r2
timeout 900 python bench.py > gpurun_out/r2/bench_n1_final.json 2> gpurun_out/r2/bench_n1_final.err; echo "bench1 rc=$?"
tail -2 gpurun_out/r2/bench_n1_final.err
