/* oracle/opcount/counted.h -- TEST INFRASTRUCTURE. Instrumented scalar types for counting the floating-point operations the
 * flat oracle executes ("replace the ~ constants with an instrumented op count from the flat oracle").
 * The oracle's C sources are compiled UNCHANGED as C++ with `double` / `float` macro-replaced by these classes
 * (oracle/opcount/Makefile); every arithmetic operator and math call bumps a thread-local counter. */
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cmath>

struct OpCounts { uint64_t add, mul, div, sqrt_, trig, log_, cmp, other; };
extern thread_local OpCounts g_ops;

template <class T> struct Counted {
    T v;
    Counted() = default;
    Counted(T x) : v(x) {}
    template <class U> Counted(const Counted<U> &o) : v((T)o.v) {}
    Counted(int x) : v((T)x) {}
    Counted(unsigned x) : v((T)x) {}
    Counted(long x) : v((T)x) {}
    Counted(unsigned long x) : v((T)x) {}
    Counted(long long x) : v((T)x) {}
    Counted(typename std::conditional<std::is_same<T, double>::value, float, double>::type x) : v((T)x) {}
    explicit operator int() const { return (int)v; }
    explicit operator long() const { return (long)v; }
    explicit operator unsigned() const { return (unsigned)v; }
    explicit operator bool() const { return v != 0; }
    Counted &operator+=(Counted o) { g_ops.add++; v += o.v; return *this; }
    Counted &operator-=(Counted o) { g_ops.add++; v -= o.v; return *this; }
    Counted &operator*=(Counted o) { g_ops.mul++; v *= o.v; return *this; }
    Counted &operator/=(Counted o) { g_ops.div++; v /= o.v; return *this; }
    Counted operator-() const { return Counted(-v); }
    Counted operator+() const { return *this; }
};
#define CBIN(op, field)                                                                                              \
    template <class T> inline Counted<T> operator op(Counted<T> a, Counted<T> b) { g_ops.field++; return Counted<T>(a.v op b.v); } \
    template <class T, class U, class = typename std::enable_if<std::is_arithmetic<U>::value>::type>                   \
    inline Counted<T> operator op(Counted<T> a, U b) { g_ops.field++; return Counted<T>(a.v op (T)b); }                 \
    template <class T, class U, class = typename std::enable_if<std::is_arithmetic<U>::value>::type>                   \
    inline Counted<T> operator op(U a, Counted<T> b) { g_ops.field++; return Counted<T>((T)a op b.v); }                 \
    inline Counted<double> operator op(Counted<double> a, Counted<float> b) { g_ops.field++; return Counted<double>(a.v op (double)b.v); } \
    inline Counted<double> operator op(Counted<float> a, Counted<double> b) { g_ops.field++; return Counted<double>((double)a.v op b.v); }
CBIN(+, add) CBIN(-, add) CBIN(*, mul) CBIN(/, div)
#undef CBIN
#define CCMP(op)                                                                                                     \
    template <class T> inline bool operator op(Counted<T> a, Counted<T> b) { g_ops.cmp++; return a.v op b.v; }          \
    template <class T, class U, class = typename std::enable_if<std::is_arithmetic<U>::value>::type>                   \
    inline bool operator op(Counted<T> a, U b) { g_ops.cmp++; return a.v op (T)b; }                                     \
    template <class T, class U, class = typename std::enable_if<std::is_arithmetic<U>::value>::type>                   \
    inline bool operator op(U a, Counted<T> b) { g_ops.cmp++; return (T)a op b.v; }                                     \
    inline bool operator op(Counted<double> a, Counted<float> b) { g_ops.cmp++; return a.v op (double)b.v; }            \
    inline bool operator op(Counted<float> a, Counted<double> b) { g_ops.cmp++; return (double)a.v op b.v; }
CCMP(<) CCMP(>) CCMP(<=) CCMP(>=) CCMP(==) CCMP(!=)
#undef CCMP
#define CFN1(name, field) \
    template <class T> inline Counted<T> c_##name(Counted<T> a) { g_ops.field++; return Counted<T>((T)std::name(a.v)); }
CFN1(sqrt, sqrt_) CFN1(sin, trig) CFN1(cos, trig) CFN1(tan, trig) CFN1(log, log_) CFN1(fabs, other) CFN1(floor, other) CFN1(ceil, other) CFN1(rint, other)
#undef CFN1
#define CFN1P(name) inline double c_##name(double a) { return std::name(a); } inline double c_##name(int a) { return std::name((double)a); }
CFN1P(sqrt) CFN1P(sin) CFN1P(cos) CFN1P(tan) CFN1P(log) CFN1P(fabs) CFN1P(floor) CFN1P(ceil) CFN1P(rint)
#undef CFN1P
template <class T> inline Counted<T> c_atan2(Counted<T> a, Counted<T> b) { g_ops.trig++; return Counted<T>((T)std::atan2(a.v, b.v)); }
template <class T> inline Counted<T> c_fmod(Counted<T> a, Counted<T> b) { g_ops.div++; return Counted<T>((T)std::fmod(a.v, b.v)); }
template <class T, class U> inline Counted<T> c_fmod(Counted<T> a, U b) { g_ops.div++; return Counted<T>((T)std::fmod(a.v, (T)b)); }
template <class T> inline Counted<T> c_fmin(Counted<T> a, Counted<T> b) { g_ops.cmp++; return Counted<T>(std::fmin(a.v, b.v)); }
template <class T> inline Counted<T> c_fmax(Counted<T> a, Counted<T> b) { g_ops.cmp++; return Counted<T>(std::fmax(a.v, b.v)); }
template <class T, class U> inline Counted<T> c_fmin(Counted<T> a, U b) { g_ops.cmp++; return Counted<T>(std::fmin(a.v, (T)b)); }
template <class T, class U> inline Counted<T> c_fmax(Counted<T> a, U b) { g_ops.cmp++; return Counted<T>(std::fmax(a.v, (T)b)); }
template <class T, class U, class = typename std::enable_if<std::is_arithmetic<U>::value>::type> inline Counted<T> c_fmin(U a, Counted<T> b) { g_ops.cmp++; return Counted<T>(std::fmin((T)a, b.v)); }
template <class T, class U, class = typename std::enable_if<std::is_arithmetic<U>::value>::type> inline Counted<T> c_fmax(U a, Counted<T> b) { g_ops.cmp++; return Counted<T>(std::fmax((T)a, b.v)); }
template <class T> inline bool c_isfinite(Counted<T> a) { return std::isfinite(a.v); }
typedef Counted<double> cdouble;
typedef Counted<float> cfloat;
