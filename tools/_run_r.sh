for v in f0 f1 f2 f4 f8 f15; do echo "== $v"; H1V2_LIB=build/variants/lib_$v.so python tools/quick_gpu_cat.py | grep 32768; done
