set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_a.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_tests_a.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_bench_a.err | cut -c1-300
