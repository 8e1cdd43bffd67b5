// 2^x evaluators shared by the dust-spectra kernel (sed_dust.cu) and the peak micro-benchmarks (peaks.cu).
#pragma once

__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float2 ex2_mufu2(float2 x) { return make_float2(ex2a(x.x), ex2a(x.y)); }

// 2^x for x <= 0 on the FMA/ALU pipes (packed FP32): the XU pipe (16 ex2/clk/SM) is this kernel's
// roof, so half of the exponentials are evaluated here instead.  Round-to-nearest split x = n + f
// via the 1.5*2^23 magic constant, degree-5 minimax polynomial of 2^f on [-0.5, 0.5] (max relative
// error 2.3e-7 in FP32 Horner form, the same order as ex2.approx), exponent add on the bit pattern.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    const float kMagic = 12582912.f;
    x.x = fmaxf(x.x, -125.f);
    x.y = fmaxf(x.y, -125.f);
    const float2 r = __fadd2_rn(x, make_float2(kMagic, kMagic));
    const float2 rn = __fadd2_rn(r, make_float2(-kMagic, -kMagic));
    const float2 f = __ffma2_rn(rn, make_float2(-1.f, -1.f), x);
    float2 q = make_float2(0.0013276471518578379f, 0.0013276471518578379f);
    q = __ffma2_rn(q, f, make_float2(0.009675541330665999f, 0.009675541330665999f));
    q = __ffma2_rn(q, f, make_float2(0.05550713274963921f, 0.05550713274963921f));
    q = __ffma2_rn(q, f, make_float2(0.24022119723934243f, 0.24022119723934243f));
    q = __ffma2_rn(q, f, make_float2(0.6931469670638628f, 0.6931469670638628f));
    q = __ffma2_rn(q, f, make_float2(1.0000000716546567f, 1.0000000716546567f));
    q.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(r.x) << 23));
    q.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(r.y) << 23));
    return q;
}

