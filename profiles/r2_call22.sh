set -x
timeout 1200 python profiles/k1_variants.py build/variants/k1_s5.so build/variants/k1_s4.so build/variants/k1_s6.so build/variants/k1_s5j2.so build/variants/k1_s4j2.so build/variants/k1_s6j2.so build/variants/k1_s5j3.so build/variants/k1_s5j2a3.so build/variants/k1_s5j2a4.so
