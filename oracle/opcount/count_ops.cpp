// oracle/opcount/count_ops.cpp -- TEST INFRASTRUCTURE. The flat oracle (../rvo2_ref.c, ../mpc_ref.c), compiled unchanged
// with instrumented scalars, so that the floating-point work per ORCA solve / per MPC damped solve can be COUNTED instead
// of estimated. Built by oracle/opcount/Makefile into libopcount.so; driven by oracle/opcount/count.py,
// which writes profiles/roofline.json.
#include "counted.h"
thread_local OpCounts g_ops = {0, 0, 0, 0, 0, 0, 0, 0};
extern "C" void opcount_get(uint64_t out[8])
{
    out[0] = g_ops.add; out[1] = g_ops.mul; out[2] = g_ops.div; out[3] = g_ops.sqrt_; out[4] = g_ops.trig; out[5] = g_ops.log_;
    out[6] = g_ops.cmp; out[7] = g_ops.other;
}
extern "C" void opcount_reset(void) { memset(&g_ops, 0, sizeof g_ops); }

#define _Thread_local thread_local
#define double cdouble
#define float cfloat
#define sqrt c_sqrt
#define sqrtf c_sqrt
#define sin c_sin
#define cos c_cos
#define tan c_tan
#define log c_log
#define fabs c_fabs
#define fabsf c_fabs
#define floor c_floor
#define ceil c_ceil
#define atan2 c_atan2
#define fmin c_fmin
#define fminf c_fmin
#define fmax c_fmax
#define fmaxf c_fmax
#define fmod c_fmod
extern "C" {
#include "../rvo2_ref.c"
#include "../mpc_ref.c"
}
