// counted_double.h -- TEST INFRASTRUCTURE (oracle/Makefile, target libh1v2_oracle_counted.so).
// Counting scalar: h1v2_oracle.c is compiled as C++ with `double` replaced by this type, so every double-precision
// arithmetic operation of the oracle increments a thread-local counter; results are bit-identical to the plain build
// (tests/test_flopcount.py).  The count for the double-support standing state is the binding FLOP figure of the FP32
// roofline, frozen in profiles/roofline.json by oracle/flopcount/count.py.
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
struct FlopCount { unsigned long long add, mul, div, fma, sqrt, trans, cmp; };
extern thread_local FlopCount g_fc;
struct cdouble {
  double v;
  cdouble() = default;
  constexpr cdouble(double x) : v(x) {}
  constexpr cdouble(float x) : v(x) {}
  constexpr cdouble(int x) : v(x) {}
  constexpr cdouble(long x) : v((double)x) {}
  constexpr cdouble(long long x) : v((double)x) {}
  constexpr cdouble(unsigned x) : v(x) {}
  explicit operator float() const { return (float)v; }
  explicit operator int() const { return (int)v; }
  explicit operator long() const { return (long)v; }
  explicit operator bool() const { return v != 0; }
  cdouble operator-() const { return cdouble(-v); }
  cdouble& operator+=(cdouble o) { g_fc.add++; v += o.v; return *this; }
  cdouble& operator-=(cdouble o) { g_fc.add++; v -= o.v; return *this; }
  cdouble& operator*=(cdouble o) { g_fc.mul++; v *= o.v; return *this; }
  cdouble& operator/=(cdouble o) { g_fc.div++; v /= o.v; return *this; }
};
#define CD_BIN(op, ctr) \
  inline cdouble operator op(cdouble a, cdouble b) { g_fc.ctr++; return cdouble(a.v op b.v); } \
  inline cdouble operator op(cdouble a, double b) { g_fc.ctr++; return cdouble(a.v op b); } \
  inline cdouble operator op(double a, cdouble b) { g_fc.ctr++; return cdouble(a op b.v); } \
  inline cdouble operator op(cdouble a, float b) { g_fc.ctr++; return cdouble(a.v op b); } \
  inline cdouble operator op(float a, cdouble b) { g_fc.ctr++; return cdouble(a op b.v); } \
  inline cdouble operator op(cdouble a, int b) { g_fc.ctr++; return cdouble(a.v op b); } \
  inline cdouble operator op(int a, cdouble b) { g_fc.ctr++; return cdouble(a op b.v); }
CD_BIN(+, add) CD_BIN(-, add) CD_BIN(*, mul) CD_BIN(/, div)
#define CD_CMP(op) \
  inline bool operator op(cdouble a, cdouble b) { g_fc.cmp++; return a.v op b.v; } \
  inline bool operator op(cdouble a, double b) { g_fc.cmp++; return a.v op b; } \
  inline bool operator op(double a, cdouble b) { g_fc.cmp++; return a op b.v; } \
  inline bool operator op(cdouble a, float b) { g_fc.cmp++; return a.v op b; } \
  inline bool operator op(float a, cdouble b) { g_fc.cmp++; return a op b.v; } \
  inline bool operator op(cdouble a, int b) { g_fc.cmp++; return a.v op b; } \
  inline bool operator op(int a, cdouble b) { g_fc.cmp++; return a op b.v; }
CD_CMP(<) CD_CMP(>) CD_CMP(<=) CD_CMP(>=) CD_CMP(==) CD_CMP(!=)
inline cdouble sqrt(cdouble a) { g_fc.sqrt++; return cdouble(sqrt(a.v)); }
inline cdouble fabs(cdouble a) { return cdouble(fabs(a.v)); }
inline cdouble fmax(cdouble a, cdouble b) { g_fc.cmp++; return cdouble(fmax(a.v, b.v)); }
inline cdouble fmin(cdouble a, cdouble b) { g_fc.cmp++; return cdouble(fmin(a.v, b.v)); }
inline cdouble fmax(cdouble a, double b) { g_fc.cmp++; return cdouble(fmax(a.v, b)); }
inline cdouble fmin(cdouble a, double b) { g_fc.cmp++; return cdouble(fmin(a.v, b)); }
inline cdouble fmax(double a, cdouble b) { g_fc.cmp++; return cdouble(fmax(a, b.v)); }
inline cdouble fmin(double a, cdouble b) { g_fc.cmp++; return cdouble(fmin(a, b.v)); }
#define CD_TR(fn) inline cdouble fn(cdouble a) { g_fc.trans++; return cdouble(fn(a.v)); }
CD_TR(sin) CD_TR(cos) CD_TR(exp) CD_TR(floor) CD_TR(ceil) CD_TR(acos) CD_TR(asin) CD_TR(tan) CD_TR(log)
inline cdouble atan2(cdouble a, cdouble b) { g_fc.trans++; return cdouble(atan2(a.v, b.v)); }
inline cdouble pow(cdouble a, cdouble b) { g_fc.trans++; return cdouble(pow(a.v, b.v)); }
inline cdouble pow(cdouble a, double b) { g_fc.trans++; return cdouble(pow(a.v, b)); }
inline cdouble fma(cdouble a, cdouble b, cdouble c) { g_fc.fma++; return cdouble(fma(a.v, b.v, c.v)); }
inline bool isfinite(cdouble a) { return std::isfinite(a.v); }
inline bool isnan(cdouble a) { return std::isnan(a.v); }
#define double cdouble
#define _Thread_local thread_local
