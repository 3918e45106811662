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

#include <type_traits>

namespace fewbit {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSubtile = 256;  // elements per warp per round of 128-bit loads

// ---------------------------------------------------------------- global memory I/O ----

// Tuning knobs (benchmarks/sweep.py builds variants with -D...): cache policy of the streaming
// loads / stores (measured on B200: no effect, profiles/r01_sweep_a.txt).
#ifndef FEWBIT_LD_MODE
#define FEWBIT_LD_MODE 1  // 0: ld.global   1: ld.global.L1::no_allocate   2: ld.global.cs
#endif
#ifndef FEWBIT_PREFETCH
// Input prefetch of the math-heavy forward kernels.  0: load a tile, compute it.
// 1: software pipeline through registers.  2: streamed half tiles (ForwardStream).
#define FEWBIT_PREFETCH 2
#endif
#ifndef FEWBIT_STREAM_MODE
// How a streamed kernel gets its input early.  0: cp.async ring in shared memory.  1: bulk
// prefetch into L2 two halves ahead, plain loads at the point of use.  2 (default): per kernel --
// the ring decouples the load latency completely and wins wherever instruction issue is the
// limit (fp32, bf16 up to 4 bits: 3-bit GELU bf16 87 % against 80 %); from 5 bits on the bf16
// kernels are bound by shared-memory wavefronts (ncu: that pipe 70-80 % busy, table gathers at
// 3.3 wavefronts each), the ring's round trip through shared memory is a quarter of those, and
// the L2 prefetch wins (7-bit hardswish 85 % against 82 %).  profiles/r02_stream_modes.txt.
#define FEWBIT_STREAM_MODE 2
#endif
#ifndef FEWBIT_BACKWARD_PIPE
#define FEWBIT_BACKWARD_PIPE 0   // backward kernels: request the next tile (ordered loads) before computing this one
#endif
#ifndef FEWBIT_ORDERED_PIPE
#define FEWBIT_ORDERED_PIPE 0    // forward kernels that do not stream: software pipeline of ordered register loads
#endif
#ifndef FEWBIT_WHOLE_TILES_8
#define FEWBIT_WHOLE_TILES_8 0   // bf16 8-bit streamed kernels: deal whole tiles (one flush per tile) instead of halves
#endif
template <typename T, int B> constexpr bool whole_tile_units() { return FEWBIT_WHOLE_TILES_8 && sizeof(T) == 2 && B == 8; }
#ifndef FEWBIT_L2_PIPELINE
#define FEWBIT_L2_PIPELINE 1   // L2-prefetch mode: load the next half's registers before computing this one
#endif
#ifndef FEWBIT_L2_AHEAD
#define FEWBIT_L2_AHEAD 3      // halves between the L2 request and the computation (pipelined L2 mode)
#endif
#ifndef FEWBIT_STREAM_L2_MAXBITS
#define FEWBIT_STREAM_L2_MAXBITS 8
#endif
template <typename T, int B> constexpr int stream_mode() {
    return FEWBIT_STREAM_MODE != 2 ? FEWBIT_STREAM_MODE : (sizeof(T) == 2 && B >= 5 && B <= FEWBIT_STREAM_L2_MAXBITS ? 1 : 0);
}
#ifndef FEWBIT_ST_MODE
#define FEWBIT_ST_MODE 1  // 0: st.global   1: st.global.L1::no_allocate   2: st.global.cs
#endif

// Not `volatile`: a load is a pure function of its address as far as these kernels go (every
// element is read once, before its own in-place overwrite, which depends on the loaded value), and
// the scheduler must be free to hoist the prefetching loads above earlier stores.
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    // Plain (coherent) load: x and y may alias, so the read-only .nc path is off limits.
#if FEWBIT_LD_MODE == 0
    asm("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
#elif FEWBIT_LD_MODE == 1
    asm("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
#else
    asm("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];"
#endif
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// The same load as an ordered statement: it stays where it is written relative to the other volatile
// statements (the shared-memory and global stores of the half computed meanwhile), which is what a
// load issued one half AHEAD of its use needs -- left to itself the compiler sinks it to the use.
__device__ __forceinline__ uint4 ldg_stream_ordered(const uint4 *p) {
    uint4 r;
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
    static constexpr int kVectors = 64;   // 128-bit vectors per subtile
    static __device__ __forceinline__ void load(const float *base, int lane, float (&v)[8]) {
        const uint4 *p = reinterpret_cast<const uint4 *>(base);
        uint4 a = ldg_stream(p + lane), b = ldg_stream(p + 32 + lane);
        v[0] = __uint_as_float(a.x), v[1] = __uint_as_float(a.y);
        v[2] = __uint_as_float(a.z), v[3] = __uint_as_float(a.w);
        v[4] = __uint_as_float(b.x), v[5] = __uint_as_float(b.y);
        v[6] = __uint_as_float(b.z), v[7] = __uint_as_float(b.w);
    }
    // the same through this lane's pointer (first vector of the subtile + lane), and split into
    // the load proper and the conversion (the pipelined loop keeps raw vectors in flight)
    struct Raw {
        uint4 a, b;
    };
    static __device__ __forceinline__ Raw fetch(const uint4 *p) { return Raw{ldg_stream(p), ldg_stream(p + 32)}; }
    static __device__ __forceinline__ Raw fetch_ordered(const uint4 *p) { return Raw{ldg_stream_ordered(p), ldg_stream_ordered(p + 32)}; }
    static __device__ __forceinline__ void widen(const Raw &r, float (&v)[8]) {
        v[0] = __uint_as_float(r.a.x), v[1] = __uint_as_float(r.a.y);
        v[2] = __uint_as_float(r.a.z), v[3] = __uint_as_float(r.a.w);
        v[4] = __uint_as_float(r.b.x), v[5] = __uint_as_float(r.b.y);
        v[6] = __uint_as_float(r.b.z), v[7] = __uint_as_float(r.b.w);
    }
    static __device__ __forceinline__ void load_at(const uint4 *p, float (&v)[8]) { widen(fetch(p), v); }
    static __device__ __forceinline__ void store_at(uint4 *p, const float (&v)[8]) {
        stg_stream(p, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]),
                                 __float_as_uint(v[2]), __float_as_uint(v[3])));
        stg_stream(p + 32, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]),
                                      __float_as_uint(v[6]), __float_as_uint(v[7])));
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
    static constexpr int kVectors = 32;
    static __device__ __forceinline__ void load(const __nv_bfloat16 *base, int lane,
                                                float (&v)[8]) {
        uint4 a = ldg_stream(reinterpret_cast<const uint4 *>(base) + lane);
        v[0] = bf16_lo(a.x), v[1] = bf16_hi(a.x), v[2] = bf16_lo(a.y), v[3] = bf16_hi(a.y);
        v[4] = bf16_lo(a.z), v[5] = bf16_hi(a.z), v[6] = bf16_lo(a.w), v[7] = bf16_hi(a.w);
    }
    struct Raw {
        uint4 a;
    };
    static __device__ __forceinline__ Raw fetch(const uint4 *p) { return Raw{ldg_stream(p)}; }
    static __device__ __forceinline__ Raw fetch_ordered(const uint4 *p) { return Raw{ldg_stream_ordered(p)}; }
    static __device__ __forceinline__ void widen(const Raw &r, float (&v)[8]) {
        v[0] = bf16_lo(r.a.x), v[1] = bf16_hi(r.a.x), v[2] = bf16_lo(r.a.y), v[3] = bf16_hi(r.a.y);
        v[4] = bf16_lo(r.a.z), v[5] = bf16_hi(r.a.z), v[6] = bf16_lo(r.a.w), v[7] = bf16_hi(r.a.w);
    }
    static __device__ __forceinline__ void load_at(const uint4 *p, float (&v)[8]) { widen(fetch(p), v); }
    static __device__ __forceinline__ void store_at(uint4 *p, const float (&v)[8]) {
        stg_stream(p, make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                 pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
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

// Four codes -> one 4*B-bit field.  Written as multiply-adds so that the shifts can go to the
// FMA pipe (IMAD) instead of the ALU pipe, which is the busy one in the forward kernels.
template <int B>
__device__ __forceinline__ uint32_t pack4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    return c0 + c1 * (1u << B) + c2 * (1u << (2 * B)) + c3 * (1u << (3 * B));
}

// A forward op hands over the codes of its 8 values as two 4*B-bit halves (values 0..3, 4..7).
// This lane's octet (B bytes, stream order) and its index 0..31 within the subtile:
//   bf16: the lane's 8 values are octet `lane`.
//   fp32: values [4l, 4l+4) are half (l & 1) of octet l >> 1, values [128+4l, 128+4l+4) are half
//         (l & 1) of octet 16 + (l >> 1).  Even lanes assemble octet l >> 1 (they need the odd
//         neighbour's first half), odd lanes octet 16 + (l >> 1) (the even neighbour's second
//         half): one shuffle.
template <typename T, int B>
__device__ __forceinline__ uint64_t lane_octet(int lane, uint32_t first, uint32_t second) {
    uint32_t low = first, high = second;
    if constexpr (sizeof(T) == 4) {
        const bool odd = lane & 1;
        const uint32_t theirs = __shfl_xor_sync(0xffffffffu, odd ? first : second, 1);
        low = odd ? theirs : first, high = odd ? second : theirs;
    }
    if constexpr (B <= 4) return (uint64_t)(low | (high << ((4 * B) & 31)));
    return (uint64_t)low | ((uint64_t)high << (4 * B));
}
template <typename T> __device__ __forceinline__ int octet_index(int lane) {
    return sizeof(T) == 2 ? lane : (lane >> 1) + ((lane & 1) ? 16 : 0);
}

// Shared-memory accesses of the staging code name their location as a 32-bit shared-space address
// plus a compile-time offset, so that every access is one instruction with an immediate offset off
// a register that is set up once per kernel (left to pointer arithmetic the compiler recomputed
// warp, lane and strip addresses for every subtile: ~12 of ~170 instructions).
template <int kOffset> __device__ __forceinline__ void sts8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0+%2], %1;" ::"r"(a), "r"(v), "n"(kOffset) : "memory");
}
template <int kOffset> __device__ __forceinline__ void sts16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0+%2], %1;" ::"r"(a), "r"(v), "n"(kOffset) : "memory");
}
template <int kOffset> __device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0+%2], %1;" ::"r"(a), "r"(v), "n"(kOffset) : "memory");
}
template <int kOffset> __device__ __forceinline__ void sts64(uint32_t a, uint32_t lo, uint32_t hi) {
    asm volatile("st.shared.v2.u32 [%0+%3], {%1, %2};" ::"r"(a), "r"(lo), "r"(hi), "n"(kOffset) : "memory");
}
template <int kOffset> __device__ __forceinline__ uint64_t lds64(uint32_t a) {
    uint32_t lo, hi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(lo), "=r"(hi) : "r"(a), "n"(kOffset) : "memory");
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
template <int kOffset> __device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a), "n"(kOffset) : "memory");
    return r;
}

// Asynchronous 16-byte copy global -> shared (LDGSTS, L1 bypassed): no register holds the data in
// flight, so nothing tempts the scheduler to delay the load until just before its use.
template <int kOffset> __device__ __forceinline__ void copy_async16(uint32_t dst, const uint4 *src) {
    asm volatile("cp.async.cg.shared.global [%0+%2], [%1], 16;" ::"r"(dst), "l"(src), "n"(kOffset) : "memory");
}
__device__ __forceinline__ void copy_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending> __device__ __forceinline__ void copy_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}
// Ask for `bytes` (a multiple of 16) at `src` (16-byte aligned) to be brought into L2; one lane
// speaks for the warp.
__device__ __forceinline__ void prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// Write the low B bytes of `octet` to the (B-byte aligned) shared-memory address a + kOffset.
template <int B, int kOffset> __device__ __forceinline__ void put_octet(uint32_t a, uint64_t octet) {
    const uint32_t lo = (uint32_t)octet, hi = (uint32_t)(octet >> 32);
    if constexpr (B == 8) {
        sts64<kOffset>(a, lo, hi);
    } else if constexpr (B == 4) {
        sts32<kOffset>(a, lo);
    } else if constexpr (B == 2) {
        sts16<kOffset>(a, lo);
    } else if constexpr (B == 6) {
        sts16<kOffset>(a, lo), sts16<kOffset + 2>(a, lo >> 16), sts16<kOffset + 4>(a, hi);
    } else {   // 1, 3, 5, 7
        sts8<kOffset>(a, lo);
        if constexpr (B >= 3) sts8<kOffset + 1>(a, lo >> 8), sts8<kOffset + 2>(a, lo >> 16);
        if constexpr (B >= 5) sts8<kOffset + 3>(a, lo >> 24), sts8<kOffset + 4>(a, hi);
        if constexpr (B >= 7) sts8<kOffset + 5>(a, hi >> 8), sts8<kOffset + 6>(a, hi >> 16);
    }
}

// Staging of the packed bytes of one warp tile (U subtiles) on their way to global memory.
//
// Direct (B = 1, 2, 3, 4, 8, or any tile that is not four subtiles): every lane drops its octet at
// its stream position in the strip -- one store for B = 1, 2, 4, 8, three byte stores for B = 3 --
// and the strip leaves with 128-bit coalesced stores.
//
// Transposed (B = 5, 6, 7 with four subtiles): B byte-or-halfword stores per octet, 2-way bank
// conflicted, were what bounded those kernels (shared-memory wavefronts, not HBM).  Instead the
// four subtiles' octets are parked as 8-byte slots in a layout from which lane m reads back the
// four CONSECUTIVE octets 4m..4m+3 of the tile (position of octet s: row s & 3, column s >> 2;
// rows start at `kRow`, chosen so that both the writes -- lane l of subtile u holds octet
// 32u + octet_index(l) -- and the reads are free of bank conflicts).  Four consecutive octets are
// 4B bytes = B whole words at word offset B*m of the tile: composed in registers, written as
// aligned words, and the strip leaves as before.
template <typename T, int B, int U> struct Stager {
    static constexpr bool kTransposed = (B == 5 || B == 6 || B == 7) && U == 4;
    static constexpr int kRow1 = sizeof(T) == 2 ? 36 : 34, kRow2 = 72, kRow3 = sizeof(T) == 2 ? 108 : 106;
    static constexpr int kSlots = kRow3 + 32;
    static constexpr int kPacked = U * subtile_bytes<B>();
    // +16 bytes: strip_bits() (backward) may touch one word past the last octet.
    static constexpr int kBytes = kTransposed ? kSlots * 8 : kPacked + 16;

    // Where this lane parks its octet of subtile 0 (later subtiles: compile-time offsets).
    static __device__ __forceinline__ uint32_t put_address(uint32_t strip, int lane) {
        const int i = octet_index<T>(lane);
        if constexpr (!kTransposed) return strip + B * i;
        const int row = i & 3;
        return strip + 8 * ((row == 0 ? 0 : row == 1 ? kRow1 : row == 2 ? kRow2 : kRow3) + (i >> 2));
    }
    template <int u> static __device__ __forceinline__ void put(uint32_t put_at, uint64_t octet) {
        if constexpr (!kTransposed)
            put_octet<B, u * subtile_bytes<B>()>(put_at, octet);
        else
            sts64<64 * u>(put_at, (uint32_t)octet, (uint32_t)(octet >> 32));
    }

    // After every lane has put() its U octets: move the tile's packed bytes to global memory.
    // `strip` is the warp's strip, `out` this lane's first 16-byte chunk of the tile's state.
    static __device__ __forceinline__ void flush(uint32_t strip, uint4 *out, int lane) {
        __syncwarp();
        if constexpr (kTransposed) {
            const uint32_t mine = strip + 8 * lane;
            const uint64_t o[4] = {lds64<0>(mine), lds64<8 * kRow1>(mine), lds64<8 * kRow2>(mine), lds64<8 * kRow3>(mine)};
            uint32_t w[B];  // the 4B bytes of octets 4*lane .. 4*lane+3
#pragma unroll
            for (int k = 0; k < B; ++k) {
                uint64_t acc = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {   // octet q covers stream bits [8Bq, 8B(q+1)) of the group
                    const int shift = 8 * B * q - 32 * k;
                    if (shift > -8 * B && shift < 32) acc |= shift >= 0 ? (o[q] << shift) : (o[q] >> -shift);
                }
                w[k] = (uint32_t)acc;
            }
            __syncwarp();   // every lane has read its slots: the strip can be overwritten
            const uint32_t words = strip + 4 * B * lane;
            if constexpr (B == 6) {
                sts64<0>(words, w[0], w[1]), sts64<8>(words, w[2], w[3]), sts64<16>(words, w[4], w[5]);
            } else {
                sts32<0>(words, w[0]), sts32<4>(words, w[1]), sts32<8>(words, w[2]), sts32<12>(words, w[3]);
                sts32<16>(words, w[4]);
                if constexpr (B == 7) sts32<20>(words, w[5]), sts32<24>(words, w[6]);
            }
            __syncwarp();
        }
        constexpr int kChunks = kPacked / 16;
        const uint32_t src = strip + 16 * lane;
        if constexpr (kChunks >= 32) stg_stream(out, lds128<0>(src));
        if constexpr (kChunks >= 64) stg_stream(out + 32, lds128<512>(src));
        if constexpr (kChunks >= 96) stg_stream(out + 64, lds128<1024>(src));
        if constexpr (kChunks >= 128) stg_stream(out + 96, lds128<1536>(src));
        constexpr int kFull = kChunks / 32 * 32;
        if constexpr (kChunks > kFull)
            if (lane < kChunks - kFull) stg_stream(out + kFull, lds128<16 * kFull>(src));
        __syncwarp();
    }
};

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

// One warp, `N` consecutive subtiles.  `xl` / `yl` / `out`: this lane's first 128-bit vector of
// the tile's input, output and packed state; `strip` / `put_at`: the warp's strip and this lane's
// parking address in it (shared-space addresses).
template <class Op, typename T, int N, bool kExact, int u = 0>
__device__ __forceinline__ void forward_subtiles(const Op &op, const typename Op::Scratch &scratch,
                                                 float (&v)[N][8], uint4 *yl, uint32_t put_at, int lane) {
    if constexpr (u < N) {
        uint32_t half[2];
        op.template apply<kExact>(scratch, v[u], half);
        Subtile<T>::store_at(yl + u * Subtile<T>::kVectors, v[u]);
        Stager<T, Op::kBits, N>::template put<u>(put_at, lane_octet<T, Op::kBits>(lane, half[0], half[1]));
        forward_subtiles<Op, T, N, kExact, u + 1>(op, scratch, v, yl, put_at, lane);
    }
}

template <class Op, typename T, int N, bool kExact>
__device__ __forceinline__ void forward_chunk(const Op &op, const typename Op::Scratch &scratch, const uint4 *xl,
                                              uint4 *yl, uint4 *out, uint32_t strip, uint32_t put_at, int lane) {
    float v[N][8];
#pragma unroll
    for (int u = 0; u < N; ++u) Subtile<T>::load_at(xl + u * Subtile<T>::kVectors, v[u]);
    forward_subtiles<Op, T, N, kExact>(op, scratch, v, yl, put_at, lane);
    Stager<T, Op::kBits, N>::flush(strip, out, lane);
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
// Keeps a loop-invariant address in its register: without it the compiler rebuilds warp / lane /
// strip addresses from the thread index inside the loop (the kernels are register-capped).
__device__ __forceinline__ uint32_t pinned(uint32_t v) {
    asm volatile("" : "+r"(v));
    return v;
}
template <typename P> __device__ __forceinline__ P *pinned(P *p) {
    asm volatile("" : "+l"(p));
    return p;
}

// Subtiles u0 .. u0+H-1 of a tile whose raw vectors are already in registers; their octets are
// parked at positions p0 .. p0+H-1 of stager S.
template <class Op, typename T, class S, bool kExact, int u0, int p0, int H, int h = 0>
__device__ __forceinline__ void forward_half(const Op &op, const typename Op::Scratch &scratch,
                                             const typename Subtile<T>::Raw (&raw)[H], uint4 *yl, uint32_t put_at,
                                             int lane) {
    if constexpr (h < H) {
        float v[8];
        uint32_t half[2];
        Subtile<T>::widen(raw[h], v);
        op.template apply<kExact>(scratch, v, half);
        Subtile<T>::store_at(yl + (u0 + h) * Subtile<T>::kVectors, v);
        S::template put<p0 + h>(put_at, lane_octet<T, Op::kBits>(lane, half[0], half[1]));
        forward_half<Op, T, S, kExact, u0, p0, H, h + 1>(op, scratch, raw, yl, put_at, lane);
    }
}

// The persistent loop of one warp.  kExact: the op's slow path (tables whose borders the cell
// look-up cannot separate) -- chosen once per block, outside the loop, so that the hot loop
// carries no branch.
template <class Op, typename T, int U, bool kExact>
__device__ __forceinline__ void forward_loop(const Op &op, const typename Op::Scratch &scratch, const T *x, T *y,
                                             uint8_t *state, int64_t ntiles, uint32_t strip) {
    constexpr int B = Op::kBits;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    const int64_t me = (int64_t)blockIdx.x * kWarps + warp;
    const int64_t rounds = ntiles / nwarps;
    constexpr int kVectors = U * Subtile<T>::kVectors, kChunks = U * subtile_bytes<B>() / 16;
    // full tiles: pointers advance by a constant stride, the loop carries no index arithmetic
    if (rounds > 0) {
        const uint4 *xl = pinned(reinterpret_cast<const uint4 *>(x) + me * kVectors + lane);
        uint4 *yl = pinned(reinterpret_cast<uint4 *>(y) + me * kVectors + lane);
        uint4 *out = pinned(reinterpret_cast<uint4 *>(state) + me * kChunks + lane);
        const uint32_t put_at = pinned(Stager<T, B, U>::put_address(strip, lane));
        const int64_t step = nwarps * kVectors, out_step = nwarps * kChunks;
        if constexpr (Op::kHeavy && U % 2 == 0 && (FEWBIT_PREFETCH == 1 || FEWBIT_ORDERED_PIPE)) {
            constexpr int H = U / 2, kHalf = H * Subtile<T>::kVectors;
            typename Subtile<T>::Raw first[H], second[H];
#pragma unroll
            // ordered loads: each half's registers are requested while the half before it is computed
            // (left unordered, the compiler sinks them to their first use)
            for (int h = 0; h < H; ++h) first[h] = Subtile<T>::fetch_ordered(xl + h * Subtile<T>::kVectors);
            for (int64_t r = 0; r < rounds; ++r, yl += step, out += out_step) {
#pragma unroll
                for (int h = 0; h < H; ++h) second[h] = Subtile<T>::fetch_ordered(xl + kHalf + h * Subtile<T>::kVectors);
                forward_half<Op, T, Stager<T, B, U>, kExact, 0, 0, H>(op, scratch, first, yl, put_at, lane);
                if (r + 1 < rounds) xl += step;     // the last round fetches its own first half again
#pragma unroll
                for (int h = 0; h < H; ++h) first[h] = Subtile<T>::fetch_ordered(xl + h * Subtile<T>::kVectors);
                forward_half<Op, T, Stager<T, B, U>, kExact, H, H, H>(op, scratch, second, yl, put_at, lane);
                Stager<T, B, U>::flush(strip, out, lane);
            }
        } else {
            for (int64_t r = 0; r < rounds; ++r, xl += step, yl += step, out += out_step)
                forward_chunk<Op, T, U, kExact>(op, scratch, xl, yl, out, strip, put_at, lane);
        }
    }
    // what is left over (< one tile per warp) is handed out in single subtiles
    const uint32_t put_at = Stager<T, B, 1>::put_address(strip, lane);
    for (int64_t sub = rounds * nwarps * U + me; sub < ntiles * U; sub += nwarps)
        forward_chunk<Op, T, 1, kExact>(
            op, scratch, reinterpret_cast<const uint4 *>(x) + sub * Subtile<T>::kVectors + lane,
            reinterpret_cast<uint4 *>(y) + sub * Subtile<T>::kVectors + lane,
            reinterpret_cast<uint4 *>(state) + sub * (subtile_bytes<B>() / 16) + lane, strip, put_at, lane);
}

// Which forward kernels stream their input through the cp.async ring (forward_stream below).
template <class Op, typename = void> struct wants_stream : std::true_type {};
template <class Op> struct wants_stream<Op, std::enable_if_t<!Op::kStreamInput>> : std::false_type {};
template <class Op, typename = void> struct wants_stream_f32 : std::false_type {};
template <class Op> struct wants_stream_f32<Op, std::enable_if_t<Op::kStreamF32>> : std::true_type {};
template <class Op, typename T, int U> constexpr bool streams_input() {
    return Op::kHeavy && wants_stream<Op>::value && U == 4 && FEWBIT_PREFETCH == 2 &&
           (sizeof(T) == 2 || wants_stream_f32<Op>::value);
}
// Bytes of the cp.async input ring of one CTA (dynamic shared memory; 0: kernel does not stream).
template <class Op, typename T, int U> constexpr int ring_bytes() {
    return streams_input<Op, T, U>() && stream_mode<T, Op::kBits>() == 0 ? kWarps * 3 * 2 * Subtile<T>::kVectors * 16 : 0;
}

// The math-heavy bf16 forward kernels: the same work, streamed.
//
// Input travels global -> shared by cp.async in HALF tiles (two subtiles, 1 KB), two halves ahead of
// the one being computed, through a ring of three per-warp slots; every lane copies and later reads
// back its own 16 bytes of each subtile, so no cross-lane synchronisation is involved, and a slot
// is overwritten one half after it was read, when the instructions that consumed those registers
// have long issued.  (Loading a whole tile into registers and then computing it left each warp
// with nothing in flight during its compute phase: ncu showed 4.3 of the 6.5 resident warps per
// scheduler waiting on global loads at any time, 58 % issue utilisation; loads prefetched into
// registers were sunk by the compiler to just before their use.)
//
// Work is dealt out in units of one half tile -- one tile for the 5/6/7-bit kernels, whose packed
// bytes are assembled from four subtiles -- strided over the warps of the grid (unit u belongs to
// warp u mod nwarps, so the grid still sweeps memory as one window); warps differ by at most one
// unit, i.e. 512 elements: with ~20 halves per warp on a 100 MB tensor, dealing whole tiles and
// mopping up with unpipelined single subtiles cost ~10 % in the tail.
template <class Op, typename T> struct ForwardStream {
    static constexpr int B = Op::kBits, H = 2;
    static constexpr bool kWhole = Stager<T, B, 2 * H>::kTransposed || whole_tile_units<T, B>();      // unit = tile (two halves)
    static constexpr int kParts = kWhole ? 2 : 1;
    using S = Stager<T, B, kParts * H>;
    static constexpr int kHalf = H * Subtile<T>::kVectors;                // 128-bit vectors per half
    static constexpr int kHalfChunks = H * subtile_bytes<B>() / 16;       // 16-byte chunks of packed state per half
    static constexpr int kSub = Subtile<T>::kVectors * 16;                // bytes of one subtile: 512 (bf16), 1024 (fp32)
    static constexpr uint32_t kSlot = H * kSub;
    static constexpr int kMode = stream_mode<T, B>();

    int lane;
    int64_t nwarps, me;
    int mine;                 // units of this warp (0: nothing to do)
    uint32_t ring_at;         // this lane's 16 bytes of slot 0
    const uint4 *ahead;       // first vector of the next half to request: the fetch stream runs two halves ahead
    const uint4 *now;         // first vector of the half to compute next (L2-prefetch mode)
    int to_fetch;
    uint32_t fill, slot;      // byte offsets of the slot to fill next / to read next
    int fetch_part;

    static __device__ __forceinline__ uint32_t next_slot(uint32_t v) { return v == 2 * kSlot ? 0u : v + kSlot; }

    __device__ __forceinline__ void fetch() {
        if (to_fetch > 0) {
            if constexpr (kMode == 1) {
                if (lane == 0) prefetch_l2(ahead, kSlot);          // `ahead` of lane 0 = start of the half
            } else {
                copy_async16<0>(ring_at + fill, ahead), copy_async16<kSub>(ring_at + fill, ahead + Subtile<T>::kVectors);
                if constexpr (sizeof(T) == 4)     // second 128-bit vector of each fp32 subtile
                    copy_async16<512>(ring_at + fill, ahead + 32), copy_async16<kSub + 512>(ring_at + fill, ahead + Subtile<T>::kVectors + 32);
            }
            --to_fetch;
            if (kParts == 1 || fetch_part == 1)
                ahead += nwarps * (kParts * kHalf) - (kParts - 1) * kHalf;
            else
                ahead += kHalf;
            fetch_part ^= 1;
        }
        if constexpr (kMode == 0) copy_async_commit();   // an empty group keeps the wait arithmetic uniform
    }

    // Before the operator builds its tables: the first two halves are already on their way.
    __device__ __forceinline__ void start(const T *x, int64_t nhalves, uint32_t ring) {
        lane = threadIdx.x & 31;
        nwarps = (int64_t)gridDim.x * kWarps;
        me = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
        const int64_t nunits = nhalves / kParts;
        mine = me < nunits ? (int)((nunits - me + nwarps - 1) / nwarps) : 0;
        ring_at = ring + 16 * lane;
        ahead = reinterpret_cast<const uint4 *>(x) + me * (kParts * kHalf) + lane;
        now = ahead;
        to_fetch = mine * kParts;
        fetch_part = 0;
        fill = 0, fetch();
        fill = kSlot, fetch();
        if constexpr (kMode == 1 && FEWBIT_L2_PIPELINE)              // registers run one half ahead: L2 requests three
            for (int extra = 2; extra < FEWBIT_L2_AHEAD; ++extra) fetch();
        slot = 0;
    }

    template <bool kExact>
    __device__ __forceinline__ void run(const Op &op, const typename Op::Scratch &scratch, T *y, uint8_t *state,
                                        uint32_t strip) {
        if (mine == 0) return;
        uint4 *yl = pinned(reinterpret_cast<uint4 *>(y) + me * (kParts * kHalf) + lane);
        uint4 *out = pinned(reinterpret_cast<uint4 *>(state) + me * (kParts * kHalfChunks) + lane);
        const uint32_t put_at = pinned(S::put_address(strip, lane));
        const int64_t y_step = nwarps * (kParts * kHalf), out_step = nwarps * (kParts * kHalfChunks);
        ahead = pinned(ahead);
        // L2-prefetch mode: the registers of the NEXT half are loaded (from L2: requested two halves ago)
        // before this half is computed, so the L2 latency hides behind a half's worth of arithmetic.
        typename Subtile<T>::Raw coming[H];
        if constexpr (kMode == 1 && FEWBIT_L2_PIPELINE) {
            coming[0] = Subtile<T>::fetch_ordered(now), coming[1] = Subtile<T>::fetch_ordered(now + Subtile<T>::kVectors);
            now += kParts == 1 ? y_step : kHalf;
        }
        for (int u = 0; u < mine; ++u, yl += y_step, out += out_step) {
#pragma unroll
            for (int part = 0; part < kParts; ++part) {
                typename Subtile<T>::Raw raw[H];
                if constexpr (kMode == 1 && FEWBIT_L2_PIPELINE) {
                    raw[0] = coming[0], raw[1] = coming[1];
                    if (part + 1 < kParts || u + 1 < mine) {
                        coming[0] = Subtile<T>::fetch_ordered(now), coming[1] = Subtile<T>::fetch_ordered(now + Subtile<T>::kVectors);
                        // after part 0 comes part 1 of the same unit; after the last part, the next unit
                        now += (kParts == 1 || part == 0) ? (kParts == 1 ? y_step : y_step - (kParts - 1) * kHalf) : kHalf;
                    }
                } else if constexpr (kMode == 1) {
                    // the half was requested into L2 two halves ago: these loads are L2 hits
                    raw[0] = Subtile<T>::fetch(now), raw[1] = Subtile<T>::fetch(now + Subtile<T>::kVectors);
                    now += (kParts == 1 || part == 1) ? y_step - (kParts - 1) * kHalf : kHalf;
                } else {
                    copy_async_wait<1>();
                    raw[0].a = lds128<0>(ring_at + slot), raw[1].a = lds128<kSub>(ring_at + slot);
                    if constexpr (sizeof(T) == 4)
                        raw[0].b = lds128<512>(ring_at + slot), raw[1].b = lds128<kSub + 512>(ring_at + slot);
                    // the slot read one half ago is the one to fill now (two ahead of this one, modulo 3)
                    fill = slot == 0 ? 2 * kSlot : slot - kSlot;
                    slot = next_slot(slot);
                }
                fetch();
                if (part == 0)
                    forward_half<Op, T, S, kExact, 0, 0, H>(op, scratch, raw, yl, put_at, lane);
                else
                    forward_half<Op, T, S, kExact, H, H, H>(op, scratch, raw, yl, put_at, lane);
            }
            S::flush(strip, out, lane);
        }
        if constexpr (kMode == 0) copy_async_wait<0>();
    }
};

template <class Op, typename T, int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) forward_tiles_kernel(const T *x, T *y, uint8_t *state,
                                                                int64_t ntiles, Op op) {
    constexpr int B = Op::kBits;
    constexpr bool kRing = streams_input<Op, T, U>();
    constexpr int kStaged = kRing && !Stager<T, B, U>::kTransposed && !whole_tile_units<T, B>() ? U / 2 : U;   // subtiles staged at a time
    constexpr int kStrip = Stager<T, B, kStaged>::kBytes > Stager<T, B, 1>::kBytes ? Stager<T, B, kStaged>::kBytes
                                                                                    : Stager<T, B, 1>::kBytes;
    __shared__ alignas(16) uint8_t strips[kWarps][(kStrip + 15) / 16 * 16];
    extern __shared__ __align__(16) uint8_t rings[];       // cp.async input ring: ring_bytes<Op, T, U>() at launch
    __shared__ typename Op::Scratch scratch;
    const uint32_t strip = (uint32_t)__cvta_generic_to_shared(strips[threadIdx.x >> 5]);
    if constexpr (kRing) {
        ForwardStream<Op, T> stream;
        stream.start(x, ntiles * 2, (uint32_t)__cvta_generic_to_shared(rings) +
                                        (threadIdx.x >> 5) * (uint32_t)(ring_bytes<Op, T, U>() / kWarps));
        op.prepare(scratch);       // table set-up overlaps the first loads
        if (op.exact())
            stream.template run<true>(op, scratch, y, state, strip);
        else
            stream.template run<false>(op, scratch, y, state, strip);
    } else {
        op.prepare(scratch);
        if (op.exact())
            forward_loop<Op, T, U, true>(op, scratch, x, y, state, ntiles, strip);
        else
            forward_loop<Op, T, U, false>(op, scratch, x, y, state, ntiles, strip);
    }
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
    if constexpr (FEWBIT_BACKWARD_PIPE) {
        // The next tile's gradients and packed codes are requested (ordered loads) before this tile is decoded and
        // multiplied: one tile's worth of loads is in flight during every compute phase.
        constexpr int kBytes = U * subtile_bytes<B>(), kChunks = kBytes / 16, kMine = (kChunks + 31) / 32;
        typename Subtile<T>::Raw g_next[U];
        uint4 s_next[kMine];
        auto request = [&](int64_t tile) {
            const uint4 *gp = reinterpret_cast<const uint4 *>(gout + tile * U * kSubtile) + lane;
#pragma unroll
            for (int u = 0; u < U; ++u) g_next[u] = Subtile<T>::fetch_ordered(gp + u * Subtile<T>::kVectors);
            const uint4 *packed = reinterpret_cast<const uint4 *>(state + tile * (int64_t)kBytes);
#pragma unroll
            for (int k = 0; k < kMine; ++k)
                if (lane + 32 * k < kChunks) s_next[k] = ldg_stream_ordered(packed + lane + 32 * k);
        };
        if (me < even) request(me);
        for (int64_t tile = me; tile < even; tile += nwarps) {
            typename Subtile<T>::Raw g_now[U];
            uint4 s_now[kMine];
#pragma unroll
            for (int u = 0; u < U; ++u) g_now[u] = g_next[u];
#pragma unroll
            for (int k = 0; k < kMine; ++k) s_now[k] = s_next[k];
            if (tile + nwarps < even) request(tile + nwarps);
            uint4 *dst = reinterpret_cast<uint4 *>(strip);
#pragma unroll
            for (int k = 0; k < kMine; ++k)
                if (lane + 32 * k < kChunks) dst[lane + 32 * k] = s_now[k];
            __syncwarp();
            T *dt = gin + tile * U * kSubtile;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float v[8];
                uint32_t code[8];
                Subtile<T>::widen(g_now[u], v);
                fetch_codes<T, B>(reinterpret_cast<const uint32_t *>(strip + u * subtile_bytes<B>()), lane, code);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = op.factor(code[j]) * v[j];
                Subtile<T>::store(dt + u * kSubtile, lane, v);
            }
            __syncwarp();
        }
    } else {
        for (int64_t tile = me; tile < even; tile += nwarps)
            backward_chunk<Op, T, U>(op, state, gout, gin, strip, tile * U, lane);
    }
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
        float v[8];
        uint32_t half[2];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = e0 + j < n ? to_float<T>(x[e0 + j]) : 0.0f;
        if (op.exact()) op.template apply<true>(scratch, v, half); else op.template apply<false>(scratch, v, half);
        uint64_t octet = (uint64_t)half[0] | ((uint64_t)half[1] << (4 * B));
        const int valid = n - e0 < 8 ? (int)(n - e0) : 8;   // pad bits of the last octet stay zero
        if (valid < 8) octet &= ((uint64_t)1 << (B * valid)) - 1;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (e0 + j < n) y[e0 + j] = from_float<T>(v[j]);
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
