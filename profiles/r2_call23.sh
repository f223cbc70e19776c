set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -3 gpurun_out/r2_bench.err; cat gpurun_out/r2_bench.json
