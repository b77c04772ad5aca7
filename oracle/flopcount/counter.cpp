// Flop counter behind oracle/flopcount/counted_double.h (test infrastructure, see oracle/Makefile target flopcount).
#include "counted_double.h"
#undef double
thread_local FlopCount g_fc = {0, 0, 0, 0, 0, 0, 0};
extern "C" void h1v2o_flops_reset(void) { g_fc = FlopCount{0, 0, 0, 0, 0, 0, 0}; }
// add, mul, div, fma, sqrt, transcendental, compare  (calling thread only: run the oracle with threads = 1)
extern "C" void h1v2o_flops_get(unsigned long long out[7]) {
  out[0] = g_fc.add; out[1] = g_fc.mul; out[2] = g_fc.div; out[3] = g_fc.fma; out[4] = g_fc.sqrt; out[5] = g_fc.trans; out[6] = g_fc.cmp;
}
