for v in e0_mx1 e1_mx2 e2_mx1_nofm e3_mx2_nofm; do echo "== $v"; H1V2_LIB=build/variants/lib_$v.so python tools/diag_fp64.py 8192 96 1.0 D; H1V2_LIB=build/variants/lib_$v.so SEED=7 python tools/diag_fp64.py 8192 96 1.0 D; done
python tools/time_variants.py 4096,32768 0 e
