set -x
python profiles/push_bench.py 24 20 0,1 0,0,0 2>&1 | tail -1
for v in t128c5 t128c4; do KMC_LIB=$PWD/build/variants/push_$v.so python profiles/push_bench.py 24 20 0,1 0,0,0 0,0,1024 0,0,3072 2>&1 | tail -3; done
KMC_LIB=$PWD/build/variants/push_t128c5.so timeout 600 python -m pytest tests/test_gpu_push.py -q -x -k "mvn10" 2>&1 | tail -3
