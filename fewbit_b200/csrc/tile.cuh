// fewbit_b200 -- warp-tile machinery shared by every elementwise kernel (sm_100a).
//
// Data layout.  Activations are flat arrays of T (float or bf16).  The packed state is the
// reference CPU codec's stream (fewbit/cpu/codec.h:33-57): element i owns stream bits
// [i*B, (i+1)*B), LSB first; 8 elements <-> B bytes ("octet").
//
// Work decomposition.  A *subtile* is 256 consecutive elements = what one warp covers with
// one round of 128-bit loads per lane (bf16: one LDG.128 = one octet per lane; fp32: two
// fully coalesced LDG.128 per lane, lane l holding elements [4l,4l+4) and [128+4l,128+4l+4)).
// A subtile packs to exactly 32*B bytes, i.e. B whole 32-byte DRAM sectors, so every state
// store is sector-aligned for any B.  A *warp tile* is U subtiles; its U*32*B packed bytes
// are staged in a per-warp shared-memory strip and written (forward) or read (backward)
// with 128-bit coalesced accesses.  Warps never synchronise with each other in the main
// loop (__syncwarp only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fewbit {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSubtile = 256;  // elements per warp per round of 128-bit loads

// ---------------------------------------------------------------- global memory I/O ----

// Tuning knobs (benchmarks/sweep.py builds variants with -D...): cache policy of the streaming
// loads / stores (measured on B200: no effect, profiles/r01_tuning_sweep.md).
#ifndef FEWBIT_LD_MODE
#define FEWBIT_LD_MODE 1  // 0: ld.global   1: ld.global.L1::no_allocate   2: ld.global.cs
#endif
#ifndef FEWBIT_ST_MODE
#define FEWBIT_ST_MODE 1  // 0: st.global   1: st.global.L1::no_allocate   2: st.global.cs
#endif

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    // Plain (coherent) load: x and y may alias, so the read-only .nc path is off limits.
#if FEWBIT_LD_MODE == 0
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
#elif FEWBIT_LD_MODE == 1
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
#else
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];"
#endif
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void stg_stream(uint4 *p, const uint4 &v) {
#if FEWBIT_ST_MODE == 0
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};"
#elif FEWBIT_ST_MODE == 1
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
#else
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
#endif
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    // cvt.rn.bf16x2.f32 d, a, b : a -> upper half, b -> lower half; round to nearest even.
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) {
    return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}

// One lane's share of a subtile: 8 values in registers, as floats.
//   fp32: v[0..3] = elements 4*lane + {0..3},  v[4..7] = elements 128 + 4*lane + {0..3}
//   bf16: v[0..7] = elements 8*lane + {0..7}
template <typename T> struct Subtile;

template <> struct Subtile<float> {
    static __device__ __forceinline__ void load(const float *base, int lane, float (&v)[8]) {
        const uint4 *p = reinterpret_cast<const uint4 *>(base);
        uint4 a = ldg_stream(p + lane), b = ldg_stream(p + 32 + lane);
        v[0] = __uint_as_float(a.x), v[1] = __uint_as_float(a.y);
        v[2] = __uint_as_float(a.z), v[3] = __uint_as_float(a.w);
        v[4] = __uint_as_float(b.x), v[5] = __uint_as_float(b.y);
        v[6] = __uint_as_float(b.z), v[7] = __uint_as_float(b.w);
    }
    static __device__ __forceinline__ void store(float *base, int lane, const float (&v)[8]) {
        uint4 *p = reinterpret_cast<uint4 *>(base);
        stg_stream(p + lane, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]),
                                        __float_as_uint(v[2]), __float_as_uint(v[3])));
        stg_stream(p + 32 + lane, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]),
                                             __float_as_uint(v[6]), __float_as_uint(v[7])));
    }
};

template <> struct Subtile<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *base, int lane,
                                                float (&v)[8]) {
        uint4 a = ldg_stream(reinterpret_cast<const uint4 *>(base) + lane);
        v[0] = bf16_lo(a.x), v[1] = bf16_hi(a.x), v[2] = bf16_lo(a.y), v[3] = bf16_hi(a.y);
        v[4] = bf16_lo(a.z), v[5] = bf16_hi(a.z), v[6] = bf16_lo(a.w), v[7] = bf16_hi(a.w);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *base, int lane,
                                                 const float (&v)[8]) {
        stg_stream(reinterpret_cast<uint4 *>(base) + lane,
                   make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                              pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
    }
};

// ------------------------------------------------------------ packed-strip staging ----

// Bytes of packed state per subtile / per warp tile.
template <int B> constexpr int subtile_bytes() { return 32 * B; }

// Write the low B bytes of `octet` to the (B-byte aligned) shared-memory address `p`.
template <int B> __device__ __forceinline__ void put_octet(uint8_t *p, uint64_t octet) {
    if constexpr (B == 8) {
        *reinterpret_cast<uint2 *>(p) = make_uint2((uint32_t)octet, (uint32_t)(octet >> 32));
    } else if constexpr (B == 4) {
        *reinterpret_cast<uint32_t *>(p) = (uint32_t)octet;
    } else if constexpr (B % 2 == 0) {
#pragma unroll
        for (int k = 0; k < B / 2; ++k)
            reinterpret_cast<uint16_t *>(p)[k] = (uint16_t)(octet >> (16 * k));
    } else {
#pragma unroll
        for (int k = 0; k < B; ++k) p[k] = (uint8_t)(octet >> (8 * k));
    }
}

// Forward: turn this lane's 8 codes of one subtile into its octet and park it in the strip.
//   code[j] belongs to v[j] of Subtile<T>.
// Four codes -> one 4*B-bit field.  Written as multiply-adds so that the shifts can go to the
// FMA pipe (IMAD) instead of the ALU pipe, which is the busy one in the forward kernels.
template <int B>
__device__ __forceinline__ uint32_t pack4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    return c0 + c1 * (1u << B) + c2 * (1u << (2 * B)) + c3 * (1u << (3 * B));
}

template <typename T, int B>
__device__ __forceinline__ void stage_codes(uint8_t *strip, int lane, const uint32_t (&code)[8]) {
    const uint32_t first = pack4<B>(code[0], code[1], code[2], code[3]);
    const uint32_t second = pack4<B>(code[4], code[5], code[6], code[7]);
    if constexpr (sizeof(T) == 2) {
        // bf16: the lane's 8 values are one octet.
        const uint64_t octet = B <= 4 ? (uint64_t)(first | (second << ((4 * B) & 31)))
                                      : ((uint64_t)first | ((uint64_t)second << (4 * B)));
        put_octet<B>(strip + B * lane, octet);
    } else {
        // fp32: elements [4l, 4l+4) are half (l & 1) of octet l >> 1, elements
        // [128+4l, 128+4l+4) are half (l & 1) of octet 16 + (l >> 1).  Even lanes assemble
        // octet l>>1 (they need the odd neighbour's `first`), odd lanes assemble octet
        // 16 + (l>>1) (they need the even neighbour's `second`): one shuffle.
        const bool odd = lane & 1;
        const uint32_t theirs = __shfl_xor_sync(0xffffffffu, odd ? first : second, 1);
        const uint32_t low = odd ? theirs : first, high = odd ? second : theirs;
        const uint64_t octet = B <= 4 ? (uint64_t)(low | (high << ((4 * B) & 31)))
                                      : ((uint64_t)low | ((uint64_t)high << (4 * B)));
        put_octet<B>(strip + B * ((lane >> 1) + (odd ? 16 : 0)), octet);
    }
}

// Backward: fetch `nbits` (<= 32) stream bits starting at bit `pos` of the strip.
// The strip is read as aligned 32-bit words; one word of slack past the end is required.
__device__ __forceinline__ uint32_t strip_bits(const uint32_t *strip, int pos, int nbits) {
    int w = pos >> 5, s = pos & 31;
    uint32_t lo = strip[w], hi = strip[w + 1];
    uint32_t r = __funnelshift_r(lo, hi, s);
    return nbits == 32 ? r : (r & ((1u << nbits) - 1u));
}

// Backward: this lane's 8 codes of one subtile (same element order as Subtile<T>).
template <typename T, int B>
__device__ __forceinline__ void fetch_codes(const uint32_t *strip, int lane, uint32_t (&code)[8]) {
    constexpr uint32_t mask = (1u << B) - 1u;
    if constexpr (sizeof(T) == 2) {
        // octet `lane`: stream bits [8*B*lane, 8*B*(lane+1))
        uint32_t lo = strip_bits(strip, 8 * B * lane, 4 * B);
        uint32_t hi = strip_bits(strip, 8 * B * lane + 4 * B, 4 * B);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            code[j] = (lo >> (B * j)) & mask;
            code[4 + j] = (hi >> (B * j)) & mask;
        }
    } else {
        uint32_t first = strip_bits(strip, 4 * B * lane, 4 * B);
        uint32_t second = strip_bits(strip, 4 * B * (32 + lane), 4 * B);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            code[j] = (first >> (B * j)) & mask;
            code[4 + j] = (second >> (B * j)) & mask;
        }
    }
}

// ---------------------------------------------------------------- kernel skeletons ----
//
// Op concepts
//   Forward op : static constexpr int kBits;
//                __device__ void prepare(smem scratch)   (block-wide, before the loop)
//                __device__ void apply(float (&v)[8], uint32_t (&code)[8]) const   v <- f(v)
//   Backward op: static constexpr int kBits;
//                __device__ void prepare(...)
//                __device__ float factor(uint32_t code) const     gin = factor * gout
//
// Both kernels are persistent: the grid is sized to the machine (SMs x resident CTAs) and
// warps stride over warp tiles.

template <int B, int U> struct StripStorage {
    // +16 bytes: strip_bits() may touch one word past the last octet.
    static constexpr int kBytes = U * subtile_bytes<B>() + 16;
    alignas(16) uint8_t bytes[kWarps][kBytes];
};

// One warp, `N` consecutive subtiles starting at subtile index `sub`.
template <class Op, typename T, int N>
__device__ __forceinline__ void forward_chunk(const Op &op, const T *x, T *y, uint8_t *state,
                                              uint8_t *strip, int64_t sub, int lane) {
    constexpr int B = Op::kBits;
    constexpr int kBytes = N * subtile_bytes<B>();
    const T *xt = x + sub * kSubtile;
    T *yt = y + sub * kSubtile;
    float v[N][8];
#pragma unroll
    for (int u = 0; u < N; ++u) Subtile<T>::load(xt + u * kSubtile, lane, v[u]);
#pragma unroll
    for (int u = 0; u < N; ++u) {
        uint32_t code[8];
        op.apply(v[u], code);
        Subtile<T>::store(yt + u * kSubtile, lane, v[u]);
        stage_codes<T, B>(strip + u * subtile_bytes<B>(), lane, code);
    }
    __syncwarp();
    uint4 *out = reinterpret_cast<uint4 *>(state + sub * (int64_t)subtile_bytes<B>());
    const uint4 *src = reinterpret_cast<const uint4 *>(strip);
#pragma unroll
    for (int i = lane; i < kBytes / 16; i += 32) stg_stream(out + i, src[i]);
    __syncwarp();
}

template <class Op, typename T, int N>
__device__ __forceinline__ void backward_chunk(const Op &op, const uint8_t *state, const T *gout,
                                               T *gin, uint8_t *strip, int64_t sub, int lane) {
    constexpr int B = Op::kBits;
    constexpr int kBytes = N * subtile_bytes<B>();
    const T *gt = gout + sub * kSubtile;
    T *dt = gin + sub * kSubtile;
    const uint4 *packed = reinterpret_cast<const uint4 *>(state + sub * (int64_t)subtile_bytes<B>());
    uint4 *dst = reinterpret_cast<uint4 *>(strip);
    float v[N][8];
#pragma unroll
    for (int u = 0; u < N; ++u) Subtile<T>::load(gt + u * kSubtile, lane, v[u]);
#pragma unroll
    for (int i = lane; i < kBytes / 16; i += 32) dst[i] = ldg_stream(packed + i);
    __syncwarp();
#pragma unroll
    for (int u = 0; u < N; ++u) {
        uint32_t code[8];
        fetch_codes<T, B>(reinterpret_cast<const uint32_t *>(strip + u * subtile_bytes<B>()), lane,
                          code);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[u][j] = op.factor(code[j]) * v[u][j];
        Subtile<T>::store(dt + u * kSubtile, lane, v[u]);
    }
    __syncwarp();
}

// Work split of the persistent grid: every warp takes the same number of full warp tiles
// (strided, so that the grid sweeps memory as one window); what is left over (< one tile per
// warp) is handed out in single subtiles, which bounds the imbalance at 256 elements per warp
// instead of U * 256.
template <class Op, typename T, int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) forward_tiles_kernel(const T *x, T *y, uint8_t *state,
                                                                int64_t ntiles, Op op) {
    constexpr int B = Op::kBits;
    __shared__ StripStorage<B, U> strips;
    __shared__ typename Op::Scratch scratch;
    op.prepare(scratch);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *strip = strips.bytes[warp];
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    const int64_t me = (int64_t)blockIdx.x * kWarps + warp;
    const int64_t even = ntiles / nwarps * nwarps;
    for (int64_t tile = me; tile < even; tile += nwarps)
        forward_chunk<Op, T, U>(op, x, y, state, strip, tile * U, lane);
    for (int64_t sub = even * U + me; sub < ntiles * U; sub += nwarps)
        forward_chunk<Op, T, 1>(op, x, y, state, strip, sub, lane);
}

template <class Op, typename T, int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) backward_tiles_kernel(const uint8_t *state, const T *gout,
                                                                 T *gin, int64_t ntiles, Op op) {
    constexpr int B = Op::kBits;
    __shared__ StripStorage<B, U> strips;
    __shared__ typename Op::Scratch scratch;
    op.prepare(scratch);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *strip = strips.bytes[warp];
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    const int64_t me = (int64_t)blockIdx.x * kWarps + warp;
    const int64_t even = ntiles / nwarps * nwarps;
    for (int64_t tile = me; tile < even; tile += nwarps)
        backward_chunk<Op, T, U>(op, state, gout, gin, strip, tile * U, lane);
    for (int64_t sub = even * U + me; sub < ntiles * U; sub += nwarps)
        backward_chunk<Op, T, 1>(op, state, gout, gin, strip, sub, lane);
}

// Ragged / unaligned path: one thread per octet, guarded scalar accesses.  Used for the
// last n % (U*256) elements of every tensor and for whole tensors whose pointers are not
// 16-byte aligned.  `first` is the index of the first element handled (a multiple of 8).
template <class Op, typename T>
__global__ void __launch_bounds__(kThreads) forward_ragged_kernel(const T *x, T *y, uint8_t *state,
                                                                 int64_t first, int64_t n, Op op) {
    constexpr int B = Op::kBits;
    __shared__ typename Op::Scratch scratch;
    op.prepare(scratch);
    const int64_t nbytes = (n * B + 7) / 8;
    const int64_t noctets = (n - first + 7) / 8;
    for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < noctets;
         o += (int64_t)gridDim.x * kThreads) {
        const int64_t e0 = first + 8 * o;
        uint64_t octet = 0;
        float v[8];
        uint32_t code[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = e0 + j < n ? to_float<T>(x[e0 + j]) : 0.0f;
        op.apply(v, code);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (e0 + j < n) {
                y[e0 + j] = from_float<T>(v[j]);
                octet |= (uint64_t)code[j] << (B * j);
            }
        }
        const int64_t b0 = (e0 / 8) * B;
#pragma unroll
        for (int k = 0; k < B; ++k)
            if (b0 + k < nbytes) state[b0 + k] = (uint8_t)(octet >> (8 * k));
    }
}

template <class Op, typename T>
__global__ void __launch_bounds__(kThreads) backward_ragged_kernel(const uint8_t *state, const T *gout,
                                                                  T *gin, int64_t first, int64_t n,
                                                                  Op op) {
    constexpr int B = Op::kBits;
    __shared__ typename Op::Scratch scratch;
    op.prepare(scratch);
    const int64_t nbytes = (n * B + 7) / 8;
    const int64_t noctets = (n - first + 7) / 8;
    for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < noctets;
         o += (int64_t)gridDim.x * kThreads) {
        const int64_t e0 = first + 8 * o;
        const int64_t b0 = (e0 / 8) * B;
        uint64_t octet = 0;
#pragma unroll
        for (int k = 0; k < B; ++k)
            if (b0 + k < nbytes) octet |= (uint64_t)state[b0 + k] << (8 * k);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (e0 + j < n) {
                uint32_t code = (uint32_t)(octet >> (B * j)) & ((1u << B) - 1u);
                gin[e0 + j] = from_float<T>(op.factor(code) * to_float<T>(gout[e0 + j]));
            }
        }
    }
}

}  // namespace fewbit
