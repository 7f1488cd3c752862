// fft_core.cuh -- register-level DFT butterflies and in-place shared-memory FFT passes.
//
// Everything here is __host__ __device__ so that the exact same index arithmetic and
// floating-point sequence can be executed serially on the CPU by the emulation library
// (csrc/emu.cu, used by the `-m "not gpu"` tests) and by the sm_100a kernels.
//
// Transform convention (matches np.fft / torch.fft, which the reference calls at
// TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:278): forward X[k] = sum_n x[n] e^{-2 pi i k n / P},
// unnormalised; the inverse here is the unnormalised adjoint (e^{+...}, no 1/P).
//
// A P-point line is transformed in place by 2 or 3 decimation-in-frequency passes with
// radices (R1, R2, R3).  The forward transform leaves the spectrum in digit-reversed
// POSITION order; the inverse (decimation in time, passes in reverse order) consumes exactly
// that order and returns natural order, so no reordering pass is ever needed:
//     position q = k1*(R2*R3) + k2*R3 + k3   <->   frequency k = k1 + R1*k2 + R1*R2*k3.
#pragma once
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>

#define TFC_HD __host__ __device__ __forceinline__

namespace tfcfft {

// cos / sin of 2*pi*k/32 -- the only butterfly constants radices <= 32 need.
constexpr float kCos32[32] = {
    1.f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f, 0.382683432f, 0.195090322f,
    0.f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f,
    -1.f, -0.98078528f, -0.923879533f, -0.831469612f, -0.707106781f, -0.555570233f, -0.382683432f, -0.195090322f,
    0.f, 0.195090322f, 0.382683432f, 0.555570233f, 0.707106781f, 0.831469612f, 0.923879533f, 0.98078528f};
constexpr float kSin32[32] = {
    0.f, 0.195090322f, 0.382683432f, 0.555570233f, 0.707106781f, 0.831469612f, 0.923879533f, 0.98078528f,
    1.f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f, 0.382683432f, 0.195090322f,
    0.f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f,
    -1.f, -0.98078528f, -0.923879533f, -0.831469612f, -0.707106781f, -0.555570233f, -0.382683432f, -0.195090322f};

// complex add / subtract as ONE packed f32x2 instruction (FADD2) on sm_100a: (re, im) are the two lanes
TFC_HD float2 cadd(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
TFC_HD float2 csub(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a * w
TFC_HD float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
// a * conj(w)
TFC_HD float2 cmulc(float2 a, float2 w) { return make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y); }

// ---------------------------------------------------------------------------------------------
// Packed pair arithmetic: one f32x2 register pair carries the SAME element of TWO independent
// transforms (tile A in .x, tile B in .y), so every butterfly add / multiply is one FADD2 / FMUL2 /
// FFMA2 on sm_100a -- half the issue slots of the scalar code for identical arithmetic.
// ---------------------------------------------------------------------------------------------
TFC_HD float2 p_add(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
TFC_HD float2 p_neg(float2 a) { return make_float2(-a.x, -a.y); }  // folds into the operand modifier
TFC_HD float2 p_sub(float2 a, float2 b) { return p_add(a, p_neg(b)); }
TFC_HD float2 p_mul(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
TFC_HD float2 p_fma(float2 a, float2 b, float2 c) {
#ifdef __CUDA_ARCH__
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
TFC_HD float2 p_dup(float s) { return make_float2(s, s); }

struct c2 {  // two complex numbers: (re.x, im.x) of transform A, (re.y, im.y) of transform B
    float2 re, im;
};
TFC_HD c2 make_c2(float2 re, float2 im) {
    c2 r;
    r.re = re;
    r.im = im;
    return r;
}
TFC_HD c2 cadd(c2 a, c2 b) { return make_c2(p_add(a.re, b.re), p_add(a.im, b.im)); }
TFC_HD c2 csub(c2 a, c2 b) { return make_c2(p_sub(a.re, b.re), p_sub(a.im, b.im)); }
// twiddle tables for packed data hold (wr, wr, wi, wi) so both lanes see the same factor
TFC_HD c2 cmul(c2 a, float4 w) {
    const float2 wr = make_float2(w.x, w.y), wi = make_float2(w.z, w.w);
    return make_c2(p_fma(a.re, wr, p_neg(p_mul(a.im, wi))), p_fma(a.re, wi, p_mul(a.im, wr)));
}
TFC_HD c2 cmulc(c2 a, float4 w) {
    const float2 wr = make_float2(w.x, w.y), wi = make_float2(w.z, w.w);
    return make_c2(p_fma(a.re, wr, p_mul(a.im, wi)), p_fma(a.im, wr, p_neg(p_mul(a.re, wi))));
}

template <int R, int K, bool INV>
TFC_HD c2 mul_w(c2 a) {
    constexpr int k32 = (K * (32 / R)) % 32;
    constexpr float c8 = 0.707106781186547524f;
    if constexpr (k32 == 0) {
        return a;
    } else if constexpr (k32 == 8) {
        return INV ? make_c2(p_neg(a.im), a.re) : make_c2(a.im, p_neg(a.re));
    } else if constexpr (k32 == 16) {
        return make_c2(p_neg(a.re), p_neg(a.im));
    } else if constexpr (k32 == 24) {
        return INV ? make_c2(a.im, p_neg(a.re)) : make_c2(p_neg(a.im), a.re);
    } else if constexpr (k32 == 4) {
        return INV ? make_c2(p_mul(p_sub(a.re, a.im), p_dup(c8)), p_mul(p_add(a.re, a.im), p_dup(c8)))
                   : make_c2(p_mul(p_add(a.re, a.im), p_dup(c8)), p_mul(p_sub(a.im, a.re), p_dup(c8)));
    } else if constexpr (k32 == 12) {
        return INV ? make_c2(p_mul(p_add(a.re, a.im), p_dup(-c8)), p_mul(p_sub(a.re, a.im), p_dup(c8)))
                   : make_c2(p_mul(p_sub(a.im, a.re), p_dup(c8)), p_mul(p_add(a.re, a.im), p_dup(-c8)));
    } else {
        constexpr float wr = kCos32[k32];
        constexpr float wi = INV ? kSin32[k32] : -kSin32[k32];
        return make_c2(p_fma(a.re, p_dup(wr), p_mul(a.im, p_dup(-wi))), p_fma(a.re, p_dup(wi), p_mul(a.im, p_dup(wr))));
    }
}

// storage element <-> arithmetic type
template <class E> struct Cx;
template <> struct Cx<float2> {
    using C = float2;
    using TW = float2;
    TFC_HD static C ld(const float2& e) { return e; }
    TFC_HD static float2 st(const C& c) { return c; }
};
template <> struct Cx<float4> {  // (reA, reB, imA, imB)
    using C = c2;
    using TW = float4;
    TFC_HD static C ld(const float4& e) { return make_c2(make_float2(e.x, e.y), make_float2(e.z, e.w)); }
    TFC_HD static float4 st(const C& c) { return make_float4(c.re.x, c.re.y, c.im.x, c.im.y); }
};

// a * W_R^K with W_R = e^{-2 pi i / R} (INV: e^{+2 pi i / R}); trivial factors cost no multiplies.
template <int R, int K, bool INV>
TFC_HD float2 mul_w(float2 a) {
    constexpr int k32 = (K * (32 / R)) % 32;
    constexpr float c8 = 0.707106781186547524f;
    if constexpr (k32 == 0) {
        return a;
    } else if constexpr (k32 == 8) {  // -i (fwd) / +i (inv)
        return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
    } else if constexpr (k32 == 16) {
        return make_float2(-a.x, -a.y);
    } else if constexpr (k32 == 24) {
        return INV ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
    } else if constexpr (k32 == 4) {  // (1 -+ i)/sqrt2
        return INV ? make_float2((a.x - a.y) * c8, (a.x + a.y) * c8) : make_float2((a.x + a.y) * c8, (a.y - a.x) * c8);
    } else if constexpr (k32 == 12) {  // (-1 -+ i)/sqrt2
        return INV ? make_float2(-(a.x + a.y) * c8, (a.x - a.y) * c8) : make_float2((a.y - a.x) * c8, -(a.x + a.y) * c8);
    } else {
        constexpr float wr = kCos32[k32];
        constexpr float wi = INV ? kSin32[k32] : -kSin32[k32];
        return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
    }
}

template <int R, bool INV, int K, class C>
TFC_HD void dft_combine(C* v, const C* e, const C* o) {
    if constexpr (K < R / 2) {
        const C t = mul_w<R, K, INV>(o[K]);
        v[K] = cadd(e[K], t);
        v[K + R / 2] = csub(e[K], t);
        dft_combine<R, INV, K + 1>(v, e, o);
    }
}

// In-register R-point DFT, natural order in and out (R in {1,2,4,8,16,32}).
template <int R, bool INV>
struct Dft {
    template <class C>
    TFC_HD static void run(C* v) {
        C e[R / 2], o[R / 2];
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            e[k] = v[2 * k];
            o[k] = v[2 * k + 1];
        }
        Dft<R / 2, INV>::run(e);
        Dft<R / 2, INV>::run(o);
        dft_combine<R, INV, 0>(v, e, o);
    }
};
template <bool INV>
struct Dft<1, INV> {
    template <class C>
    TFC_HD static void run(C*) {}
};

// Radix plan per line length.
template <int P> struct Plan;
template <> struct Plan<16>  { static constexpr int R1 = 4,  R2 = 4,  R3 = 1; };
template <> struct Plan<32>  { static constexpr int R1 = 8,  R2 = 4,  R3 = 1; };
template <> struct Plan<64>  { static constexpr int R1 = 8,  R2 = 8,  R3 = 1; };
template <> struct Plan<128> { static constexpr int R1 = 16, R2 = 8,  R3 = 1; };
template <> struct Plan<256> { static constexpr int R1 = 16, R2 = 16, R3 = 1; };
template <> struct Plan<512> { static constexpr int R1 = 8,  R2 = 8,  R3 = 8; };

template <int P>
TFC_HD int freq_of_pos(int q) {
    using Pl = Plan<P>;
    const int k1 = q / (Pl::R2 * Pl::R3), rem = q % (Pl::R2 * Pl::R3);
    const int k2 = rem / Pl::R3, k3 = rem % Pl::R3;
    return k1 + Pl::R1 * k2 + Pl::R1 * Pl::R2 * k3;
}
template <int P>
TFC_HD int pos_of_freq(int k) {
    using Pl = Plan<P>;
    const int k1 = k % Pl::R1, k2 = (k / Pl::R1) % Pl::R2, k3 = k / (Pl::R1 * Pl::R2);
    return k1 * (Pl::R2 * Pl::R3) + k2 * Pl::R3 + k3;
}
// position of the frequency -k (mod P)
template <int P>
TFC_HD int neg_pos(int q) { return pos_of_freq<P>((P - freq_of_pos<P>(q)) & (P - 1)); }

// Execution context: the kernels pass {threadIdx, blockDim, __syncthreads}; the CPU emulation
// passes a single serial "thread".
struct SerialCtx {
    int tid = 0, nthreads = 1;
    TFC_HD void sync() const {}
    TFC_HD void warp_sync() const {}
    TFC_HD void mark(int) const {}
};
#ifdef __CUDACC__
struct BlockCtx {
    int tid, nthreads;
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
    __device__ __forceinline__ void mark(int) const {}
};
// Block size known at compile time: the task loops get constant trip counts and unroll, so the loads of
// a thread's next task overlap the arithmetic of the current one.  `trace` (debug) records clock64 at
// stage boundaries.
template <int NT>
struct BlockCtxT {
    int tid;
    static constexpr int nthreads = NT;
    long long* trace;
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
    __device__ __forceinline__ void mark(int k) const {
        if (trace != nullptr && tid == 0) {  // global nanosecond timer: comparable across SMs
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            trace[k] = (long long)t;
        }
    }
};
#endif

// One radix-R pass over `1 << log2_lines` independent P-point lines held in memory `s`:
// element e of line l lives at s[l*ls + e*es].  L is the current DIF block length.
// tw[t] = e^{-2 pi i t / P}, t in [0, P)  (float2, or (wr,wr,wi,wi) float4 for packed data).
// LINE_FAST: consecutive threads take consecutive lines (else consecutive butterflies of a line).
template <int P, int R, int L, bool INV, bool LINE_FAST = true, class E, class Ctx>
TFC_HD void fft_pass(const Ctx& ctx, E* s, int es, int ls, int log2_lines, const typename Cx<E>::TW* tw) {
    using C = typename Cx<E>::C;
    constexpr int M = L / R;    // distance between butterfly legs (in elements)
    constexpr int JT = P / R;   // butterflies per line
    const int ntask = JT << log2_lines;
    const int lmask = (1 << log2_lines) - 1;
    for (int t = ctx.tid; t < ntask; t += ctx.nthreads) {
        const int line = LINE_FAST ? (t & lmask) : (t / JT);
        const int jj = LINE_FAST ? (t >> log2_lines) : (t % JT);
        const int blk = jj / M, j = jj % M;
        E* base = s + line * ls + (blk * L + j) * es;
        C v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = Cx<E>::ld(base[m * M * es]);
        if constexpr (!INV) {
            Dft<R, false>::run(v);
            if constexpr (M > 1) {
#pragma unroll
                for (int k = 1; k < R; ++k) v[k] = cmul(v[k], tw[(P / L) * j * k]);
            }
        } else {
            if constexpr (M > 1) {
#pragma unroll
                for (int k = 1; k < R; ++k) v[k] = cmulc(v[k], tw[(P / L) * j * k]);
            }
            Dft<R, true>::run(v);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) base[k * M * es] = Cx<E>::st(v[k]);
    }
}

// All passes of a batch of lines, each followed by a barrier.
template <int P, bool INV, class E, class Ctx>
TFC_HD void fft_lines(const Ctx& ctx, E* s, int es, int ls, int log2_lines, const typename Cx<E>::TW* tw) {
    using Pl = Plan<P>;
    constexpr int L2 = P / Pl::R1, L3 = P / (Pl::R1 * Pl::R2);
    if constexpr (!INV) {
        fft_pass<P, Pl::R1, P, false>(ctx, s, es, ls, log2_lines, tw);
        ctx.sync();
        if constexpr (Pl::R2 > 1) {
            fft_pass<P, Pl::R2, L2, false>(ctx, s, es, ls, log2_lines, tw);
            ctx.sync();
        }
        if constexpr (Pl::R3 > 1) {
            fft_pass<P, Pl::R3, L3, false>(ctx, s, es, ls, log2_lines, tw);
            ctx.sync();
        }
    } else {
        if constexpr (Pl::R3 > 1) {
            fft_pass<P, Pl::R3, L3, true>(ctx, s, es, ls, log2_lines, tw);
            ctx.sync();
        }
        if constexpr (Pl::R2 > 1) {
            fft_pass<P, Pl::R2, L2, true>(ctx, s, es, ls, log2_lines, tw);
            ctx.sync();
        }
        fft_pass<P, Pl::R1, P, true>(ctx, s, es, ls, log2_lines, tw);
        ctx.sync();
    }
}

// tw[t] = (cos 2 pi t / P, -sin 2 pi t / P)
template <int P, class Ctx>
TFC_HD void fill_twiddles(const Ctx& ctx, float2* tw) {
    for (int t = ctx.tid; t < P; t += ctx.nthreads) {
        float sn, cs;
#ifdef __CUDA_ARCH__
        sincospif(2.0f * (float)t / (float)P, &sn, &cs);
#else
        const double a = 2.0 * 3.14159265358979323846 * (double)t / (double)P;
        sn = (float)sin(a);
        cs = (float)cos(a);
#endif
        tw[t] = make_float2(cs, -sn);
    }
}

constexpr __host__ __device__ int ilog2_c(int v) { return v <= 1 ? 0 : 1 + ilog2_c(v >> 1); }

TFC_HD int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

}  // namespace tfcfft
