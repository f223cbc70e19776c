set -x
for lib in "" build/variants/bulk_nopipe.so build/variants/bulk_pipe2.so; do
  if [ -n "$lib" ]; then export KMC_LIB=$PWD/$lib; else unset KMC_LIB; fi
  timeout 200 python profiles/push_bench.py 24 10 0 2>&1 | tail -1
done
unset KMC_LIB
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5
