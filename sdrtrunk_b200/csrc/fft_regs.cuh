// In-register mixed-radix inverse DFTs (e^{+j 2 pi n k / N}, unscaled) for the channelizer's two-step FFT.
// Stockham autosort passes, decimation in frequency, over two register arrays; every index is a compile-time
// constant after unrolling, so the arrays live in registers and the twiddles become immediates.
//   y[q + s (r p + k)] = (sum_j x[q + s (p + m j)] W_r^{jk}) * W_N^{s p k},   m = N / (r s)
#pragma once
#include <cuda_runtime.h>

#include "fft_tables.cuh"

namespace rfft {

// complex add / subtract as ONE packed instruction (sm_100 add.rn.f32x2: both rails correctly rounded, i.e. the same bits as
// two scalar adds; the subtraction's negation folds into the FADD2 operand)
#ifndef SDRGPU_FFT_SCALAR
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ float2 jmul(float2 a) { return make_float2(-a.y, a.x); }   // * (+j)
// (packed: a.y (-wi, wr), then a.x (wr, wi) + that = two FFMA2 for 2 FMUL + 2 FFMA -- tried: the twiddle pairs are no longer
// 32-bit immediates but 64-bit operands fetched by LDC.64 / MOV pairs, and the kernel grew from 3592 to 3648 instructions)
__device__ __forceinline__ float2 cmul(float2 a, float wr, float wi)
{
    return make_float2(fmaf(a.x, wr, -a.y * wi), fmaf(a.x, wi, a.y * wr));
}

// a *= W_N^idx with the trivial rotations special-cased (idx is a constant after unrolling)
template <int N>
__device__ __forceinline__ float2 twiddle(float2 a, int idx)
{
    idx %= N;
    if (idx == 0) return a;
    if (4 * idx == N) return jmul(a);
    if (2 * idx == N) return make_float2(-a.x, -a.y);
    if (4 * idx == 3 * N) return make_float2(a.y, -a.x);
    return cmul(a, tw_cos<N>(idx), tw_sin<N>(idx));
}

template <int R>
__device__ __forceinline__ void butterfly(const float2 (&x)[R], float2 (&y)[R]);

template <>
__device__ __forceinline__ void butterfly<2>(const float2 (&x)[2], float2 (&y)[2])
{
    y[0] = cadd(x[0], x[1]);
    y[1] = csub(x[0], x[1]);
}

template <>
__device__ __forceinline__ void butterfly<3>(const float2 (&x)[3], float2 (&y)[3])
{
    const float sq = 0.86602540378443865f;
    const float2 t = cadd(x[1], x[2]), u = jmul(csub(x[1], x[2]));
    const float2 mm = make_float2(fmaf(-0.5f, t.x, x[0].x), fmaf(-0.5f, t.y, x[0].y));
    const float2 nn = make_float2(sq * u.x, sq * u.y);
    y[0] = cadd(x[0], t);
    y[1] = cadd(mm, nn);
    y[2] = csub(mm, nn);
}

template <>
__device__ __forceinline__ void butterfly<4>(const float2 (&x)[4], float2 (&y)[4])
{
    const float2 t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]), t2 = cadd(x[1], x[3]), t3 = jmul(csub(x[1], x[3]));
    y[0] = cadd(t0, t2);
    y[1] = cadd(t1, t3);
    y[2] = csub(t0, t2);
    y[3] = csub(t1, t3);
}

template <>
__device__ __forceinline__ void butterfly<5>(const float2 (&x)[5], float2 (&y)[5])
{
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
    const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    const float2 t1 = cadd(x[1], x[4]), t2 = cadd(x[2], x[3]), t3 = csub(x[1], x[4]), t4 = csub(x[2], x[3]);
    const float2 m1 = make_float2(fmaf(c2, t2.x, fmaf(c1, t1.x, x[0].x)), fmaf(c2, t2.y, fmaf(c1, t1.y, x[0].y)));
    const float2 m2 = make_float2(fmaf(c1, t2.x, fmaf(c2, t1.x, x[0].x)), fmaf(c1, t2.y, fmaf(c2, t1.y, x[0].y)));
    const float2 n1 = jmul(make_float2(fmaf(s2, t4.x, s1 * t3.x), fmaf(s2, t4.y, s1 * t3.y)));
    const float2 n2 = jmul(make_float2(fmaf(-s1, t4.x, s2 * t3.x), fmaf(-s1, t4.y, s2 * t3.y)));
    y[0] = make_float2(x[0].x + t1.x + t2.x, x[0].y + t1.y + t2.y);
    y[1] = cadd(m1, n1);
    y[2] = cadd(m2, n2);
    y[3] = csub(m2, n2);
    y[4] = csub(m1, n1);
}

// one Stockham pass of radix R with stride S (= product of the radices of the previous passes)
template <int N, int R, int S>
__device__ __forceinline__ void pass(const float2 (&x)[N], float2 (&y)[N])
{
    constexpr int Mm = N / (R * S);
#pragma unroll
    for (int p = 0; p < Mm; p++) {
#pragma unroll
        for (int q = 0; q < S; q++) {
            float2 in[R], out[R];
#pragma unroll
            for (int j = 0; j < R; j++) in[j] = x[q + S * (p + Mm * j)];
            butterfly<R>(in, out);
#pragma unroll
            for (int k = 0; k < R; k++) y[q + S * (R * p + k)] = twiddle<N>(out[k], S * p * k);
        }
    }
}

// N-point inverse DFT in place (natural order in, natural order out)
template <int N>
__device__ __forceinline__ void dft(float2 (&a)[N]);

#define RFFT_COPY(dst, src, n)                  \
    _Pragma("unroll") for (int i_ = 0; i_ < (n); i_++)(dst)[i_] = (src)[i_]

template <> __device__ __forceinline__ void dft<2>(float2 (&a)[2]) { float2 b[2]; pass<2, 2, 1>(a, b); RFFT_COPY(a, b, 2); }
template <> __device__ __forceinline__ void dft<3>(float2 (&a)[3]) { float2 b[3]; pass<3, 3, 1>(a, b); RFFT_COPY(a, b, 3); }
template <> __device__ __forceinline__ void dft<4>(float2 (&a)[4]) { float2 b[4]; pass<4, 4, 1>(a, b); RFFT_COPY(a, b, 4); }
template <> __device__ __forceinline__ void dft<5>(float2 (&a)[5]) { float2 b[5]; pass<5, 5, 1>(a, b); RFFT_COPY(a, b, 5); }
template <> __device__ __forceinline__ void dft<6>(float2 (&a)[6]) { float2 b[6]; pass<6, 2, 1>(a, b); pass<6, 3, 2>(b, a); }
template <> __device__ __forceinline__ void dft<8>(float2 (&a)[8]) { float2 b[8]; pass<8, 4, 1>(a, b); pass<8, 2, 4>(b, a); }
template <> __device__ __forceinline__ void dft<10>(float2 (&a)[10]) { float2 b[10]; pass<10, 2, 1>(a, b); pass<10, 5, 2>(b, a); }
template <> __device__ __forceinline__ void dft<12>(float2 (&a)[12]) { float2 b[12]; pass<12, 4, 1>(a, b); pass<12, 3, 4>(b, a); }
template <> __device__ __forceinline__ void dft<16>(float2 (&a)[16]) { float2 b[16]; pass<16, 4, 1>(a, b); pass<16, 4, 4>(b, a); }
template <> __device__ __forceinline__ void dft<20>(float2 (&a)[20]) { float2 b[20]; pass<20, 4, 1>(a, b); pass<20, 5, 4>(b, a); }
template <> __device__ __forceinline__ void dft<24>(float2 (&a)[24])
{
    float2 b[24];
    pass<24, 4, 1>(a, b);
    pass<24, 2, 4>(b, a);
    pass<24, 3, 8>(a, b);
    RFFT_COPY(a, b, 24);
}
template <> __device__ __forceinline__ void dft<25>(float2 (&a)[25]) { float2 b[25]; pass<25, 5, 1>(a, b); pass<25, 5, 5>(b, a); }
template <> __device__ __forceinline__ void dft<32>(float2 (&a)[32])
{
    float2 b[32];
    pass<32, 4, 1>(a, b);
    pass<32, 4, 4>(b, a);
    pass<32, 2, 16>(a, b);
    RFFT_COPY(a, b, 32);
}
#undef RFFT_COPY

}  // namespace rfft
