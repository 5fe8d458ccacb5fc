// Straight-line FP64 exp / sqrt for the kernel transforms.
//
// CUDA's exp() and sqrt() are accurate but, inlined 32x per thread and stage, they spend more issue slots on
// integer/uniform-register bookkeeping (rematerialised constants, slow-path branches) than on FP64 work, and K1 is
// issue-bound.  These versions assume what the kernels guarantee (finite arguments, exp argument <= 0, sqrt
// argument >= 0), take their constants from constant memory (FP64 instructions read c[bank][offset] operands
// directly) and have no branches.  Accuracy: <= 1 ulp (sqrt), <= 2 ulp (exp): far inside the 1e-9 parity gate and
// the same for every thread, so bit-equal inputs still give bit-equal outputs.
#pragma once
#include <cuda_runtime.h>

namespace skb {

__constant__ double c_exp[18] = {
    1.4426950408889634,                 // [0] log2(e)
    6755399441055744.0,                 // [1] 1.5 * 2^52: adding it rounds to the nearest integer
    -6.93147180369123816490e-01,        // [2] -ln2 (high part, 32 trailing zero bits)
    -1.90821492927058770002e-10,        // [3] -ln2 (low part)
    1.0 / 6227020800.0,                 // [4] 1/13!
    1.0 / 479001600.0,                  // [5] 1/12!
    1.0 / 39916800.0,                   // [6] 1/11!
    1.0 / 3628800.0,                    // [7] 1/10!
    1.0 / 362880.0,                     // [8] 1/9!
    1.0 / 40320.0,                      // [9] 1/8!
    1.0 / 5040.0,                       // [10] 1/7!
    1.0 / 720.0,                        // [11] 1/6!
    1.0 / 120.0,                        // [12] 1/5!
    1.0 / 24.0,                         // [13] 1/4!
    1.0 / 6.0,                          // [14] 1/3!
    0.5,                                // [15] 1/2!
    1.0,                                // [16]
    -708.0                              // [17] below this exp() is flushed to 0 (true value < 2.3e-308)
};

// exp(x) for finite x <= 0.  x = n ln2 + r, |r| <= ln2/2; exp(r) by its degree-13 Taylor polynomial (remainder
// r^14/14! < 5e-18); the 2^n scaling is an integer add on the exponent field (n >= -1022 after the flush).
__device__ __forceinline__ double exp_nonpos(double x) {
    const double fn = fma(x, c_exp[0], c_exp[1]);
    const double n = fn - c_exp[1];
    double r = fma(n, c_exp[2], x);
    r = fma(n, c_exp[3], r);
    double p = c_exp[4];
#pragma unroll
    for (int i = 5; i <= 16; ++i) p = fma(p, r, c_exp[i]);
    p = fma(p, r, c_exp[16]);                               // ... + r/1! + 1/0!
    const int ni = __double2loint(fn);                      // low word of (1.5*2^52 + n) holds n in two's complement
    const double scaled = __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
    const double y = x < c_exp[17] ? 0.0 : scaled;
    return x != x ? x : y;                                  // NaN in, NaN out (the exponent patch could hide it)
}

// sqrt(s) for finite s >= 0: hardware rsqrt seed (about 22 bits), two coupled Newton (Goldschmidt) steps and a final
// residual correction.
__device__ __forceinline__ double sqrt_nonneg(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    double g = s * y;                 // ~ sqrt(s)
    double h = 0.5 * y;               // ~ 1 / (2 sqrt(s))
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, s);
    g = fma(d, h, g);
    // s == 0 or a (flushed) denormal: the seed is inf and g is NaN -> the root is 0 to 1e-145; NaN stays NaN
    return s > 1e-290 ? g : (s != s ? s : 0.0);
}

}  // namespace skb
