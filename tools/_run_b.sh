for v in b0_f1 b1_mx; do
  echo "== $v"
  H1V2_LIB=build/variants/lib_$v.so python tools/diag_fp64.py 8192 96 1.0 D
  H1V2_LIB=build/variants/lib_$v.so SEED=5 python tools/diag_fp64.py 8192 96 0.3 D
  H1V2_LIB=build/variants/lib_$v.so DECIM=4 python tools/diag_fp64.py 8192 24 1.0 D
done
python tools/diag_fp64.py 8192 96 1.0 AB
python tools/time_variants.py 4096,32768 0 b
