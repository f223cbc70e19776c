set -x
python profiles/push_bench.py 24 20 0,1 0,0,0 2>&1 | tail -1
for v in k1 k2 k4 k8 k16 k31; do echo "== $v"; KMC_LIB=$PWD/build/variants/push_$v.so timeout 120 python profiles/push_bench.py 24 20 0,1 0,0,0 2>&1 | tail -1; done
