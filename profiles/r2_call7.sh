set -x
timeout 900 python -m pytest tests/test_gpu_rows.py tests/test_gpu_push.py -m gpu -x -q 2>&1 | tail -15
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 300 python profiles/l2_fetch_experiment.py 2>&1 | tail -4
timeout 200 python profiles/push_bench.py 24 10 0p 2>&1 | tail -1
KMC_LIB=$PWD/build/variants/push_prof.so timeout 200 python profiles/push_bench.py 24 10 0p 2>&1 | tail -2
# ncu --set full of the default fused dense-Gaussian kernel (K2G), one launch of 50 iterations
. profiles/capture_final.sh.lib
SKIP=1 KMC_TC=1 cap r2_k2g_gaussian_fused2 gaussian_fused2 python profiles/prof_run.py gaussian100d 50 0
mv gpurun_out/ncu_final_r2_k2g_gaussian_fused2.csv gpurun_out/r2_ncu_k2g_gaussian_fused2.csv
cat gpurun_out/r2_ncu_k2g_gaussian_fused2.csv | head -60
