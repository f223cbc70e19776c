set -x
for v in x2p4 x2p3 x2p2 base; do KMC_LIB=$PWD/build/variants/k3_$v.so timeout 300 python profiles/k3_variants.py 2>&1 | tail -4; done
