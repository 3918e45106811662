// fewbit_b200 -- per-element operators plugged into the warp-tile kernels of tile.cuh.
#pragma once

#include <math_constants.h>

#include <type_traits>

#include "tile.cuh"

// The bucket search below computes table indices as the bit patterns of denormal floats; with
// flush-to-zero every look-up would silently hit entry 0.
#ifdef __CUDA_FTZ
#error "fewbit_b200 must be compiled without -ftz=true / --use_fast_math (denormal arithmetic is load-bearing)"
#endif

namespace fewbit {

struct NoScratch {
    int unused;
};

// =====================================================================================
// Bucket search:  code(x) = #{ i : bounds[i] < x }   (std::lower_bound on sorted bounds;
// reference BinarySearch fewbit/cuda/codec.cu:118-131).  NaN -> 0 because every
// comparison is false, exactly as in the reference CUDA kernel.
//
//  B <= 2 : the (<= 3) bounds live in registers; a compare/select tree, B compares.
//  B >= 3 : cell look-up.  The kernels are bound by the ALU pipe (FSETP/FSEL/LOP3), not by
//           HBM, so the search is moved off it: a monotone map  cell(x) = round(sat(x*s + o)
//           * (C-1))  (two FFMAs on the FMA pipe) indexes a C = 128..2048 entry shared-memory
//           table holding  k = #{ i : cell(bounds[i]) < cell }.  Because cell() is monotone,
//           bounds in earlier cells are < x and bounds in later cells are > x, so
//           code(x) = k + (bounds[k] < x)  -- one gather, one compare -- whenever at most one
//           bound falls into each cell.  If some cell holds two or more bounds the whole
//           block falls back to an exact branch-free binary search (never the case for the
//           built-in tables or make_table(), whose borders these cell counts separate).
// Tables shorter than 2^B - 1 are padded with +inf (never counted).
// =====================================================================================

#ifndef FEWBIT_ADAPTIVE_CELLS
#define FEWBIT_ADAPTIVE_CELLS 1   // tables of 5+ bits: smallest power-of-two cell count that separates the borders
#endif
#ifndef FEWBIT_FEW_CELLS
#define FEWBIT_FEW_CELLS 0   // A/B switch: try a bank-conflict-free 32-cell table first (bf16, 3-4 bits).  Measured:
                             // no effect (3-bit GELU bf16 86 % either way, profiles/r02_stream_modes.txt) -- with the
                             // input ring those kernels are bound by instruction issue, not by the gather
#endif
template <typename T, int B, bool kInRegisters = (B <= 2)> struct Bucketizer;

template <typename T, int B> struct Bucketizer<T, B, true> {
    static constexpr int kCount = (1 << B) - 1;
    using Scratch = NoScratch;
    const T *bounds;
    int nbounds;
    float b[kCount];

    __device__ __forceinline__ void prepare(Scratch &) {
#pragma unroll
        for (int i = 0; i < kCount; ++i)
            b[i] = i < nbounds ? to_float<T>(bounds[i]) : CUDART_INF_F;
    }

    static constexpr bool kMayCrowd = false;
    template <bool>
    __device__ __forceinline__ void lookup(const Scratch &, const float (&x)[8], uint32_t (&half)[2]) const {
        uint32_t code[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if constexpr (B == 1) {
                code[j] = b[0] < x[j] ? 1u : 0u;
            } else {
                const bool p1 = b[1] < x[j];
                const bool p2 = (p1 ? b[2] : b[0]) < x[j];
                code[j] = (p1 ? 2u : 0u) | (p2 ? 1u : 0u);
            }
        }
        half[0] = pack4<B>(code[0], code[1], code[2], code[3]);
        half[1] = pack4<B>(code[4], code[5], code[6], code[7]);
    }
};

// acc += kInc if a < b: one FSETP and one predicated add (left to the compiler the same
// expression became an add, a compare and a predicated move).
template <uint32_t kInc> __device__ __forceinline__ void add_if_less(uint32_t &acc, float a, float b) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.u32 %0, %0, %3;\n\t}"
        : "+r"(acc) : "f"(a), "f"(b), "n"(kInc));
}

__device__ __forceinline__ unsigned long long pack2(float v) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
    return r;
}

// Cell index of x: sat() clamps (and sends NaN to 0); the multiply works in the denormal range,
// where the bit pattern of a float *is* its value in units of 2^-149, so the bits of the product
// are round(t * (C-1)) -- an integer straight off the FMA pipe, no conversion instruction.
// (Needs denormals: the build refuses -ftz=true / --use_fast_math, see the #error above.)
struct CellMap {
    float scale, offset;
    float last_cell;      // the number of cells - 1, as the denormal float whose bit pattern it is
    __device__ __forceinline__ void set_cells(int cells) { last_cell = __uint_as_float((uint32_t)(cells - 1)); }
    template <typename T> __device__ __forceinline__ void fit(const T *bounds, int nbounds) {
        const float lo = nbounds > 0 ? to_float<T>(bounds[0]) : 0.0f;
        const float hi = nbounds > 0 ? to_float<T>(bounds[nbounds - 1]) : 0.0f;
        const float span = hi - lo;
        if (span > 0.0f && span < CUDART_INF_F) {
            scale = 1.0f / span;
            offset = -lo * scale;
        } else {  // a single (or degenerate) border: any monotone map will do
            scale = 1.0f;
            offset = 0.5f - lo;
        }
        if (!(fabsf(offset) < CUDART_INF_F)) offset = 0.5f;
    }
    __device__ __forceinline__ uint32_t cell(float x) const {
        const float t = __saturatef(fmaf(x, scale, offset));
        return __float_as_uint(__fmul_rn(t, last_cell));
    }
    // two elements per multiply (mul.rn.f32x2): the clamp has no packed form, the product does
    __device__ __forceinline__ void cells(float x0, float x1, uint32_t &c0, uint32_t &c1) const {
        const float t0 = __saturatef(fmaf(x0, scale, offset));
        const float t1 = __saturatef(fmaf(x1, scale, offset));
        unsigned long long t2, a2;
        asm("mov.b64 %0, {%1, %2};" : "=l"(t2) : "f"(t0), "f"(t1));
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a2) : "l"(t2), "l"(pack2(last_cell)));
        asm("mov.b64 {%0, %1}, %2;" : "=r"(c0), "=r"(c1) : "l"(a2));
    }
};

// Exact branch-free binary search over `kSize - 1` sorted borders padded with +inf; out of line
// on purpose (rare: only tables whose borders the cells cannot separate; keeps the hot loop small).
// The table is named by its shared-space address: handing a generic pointer to an out-of-line
// function makes the compiler address ALL of the kernel's shared memory through a window base
// register (one extra add per look-up in the hot loop).
template <int kSize> __device__ __noinline__ uint32_t search_exact(uint32_t sorted, float x) {
    uint32_t k = 0;
    for (int step = kSize >> 1; step >= 1; step >>= 1) {
        float v;
        asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sorted + 4u * (k + step - 1)));
        if (v < x) k += step;
    }
    return k;
}

// The smallest cell count among `fewest`, 2 `fewest`, .. (below `most`) at which no two borders fall into one
// cell, else `most`; leaves `map` set to it.  Borders are sorted, so cells are monotone and only neighbours can
// clash.  One block-wide vote per candidate; nothing to vote on when fewest == most.
template <typename T>
__device__ __forceinline__ int choose_cells(CellMap &map, const T *bounds, int nbounds, int fewest, int most) {
    int cells = fewest;
    for (; cells < most; cells *= 2) {
        map.set_cells(cells);
        int clash = 0;
        for (int i = threadIdx.x; i + 1 < nbounds; i += blockDim.x)
            clash |= map.cell(to_float<T>(bounds[i])) == map.cell(to_float<T>(bounds[i + 1])) ? 1 : 0;
        if (!__syncthreads_or(clash)) return cells;
    }
    map.set_cells(most);
    return most;
}

// Borders -> cells, shared by both table layouts: fills s.bounds / s.bound_cell, then calls
// `emit(cell, k, has_border)` for every cell with k = #{borders in earlier cells}; returns
// (block-wide) whether some cell holds two or more borders.
template <int kSize, typename T, class Scratch, class Emit>
__device__ __forceinline__ bool build_cells(const CellMap &map, int cells, const T *bounds, int nbounds, Scratch &s,
                                            Emit emit) {
    __syncthreads();       // a second attempt (another cell count) overwrites what the first one read
    for (int i = threadIdx.x; i < kSize; i += blockDim.x) {
        const float v = i < nbounds ? to_float<T>(bounds[i]) : CUDART_INF_F;
        s.bounds[i] = v;
        s.bound_cell[i] = i < nbounds ? (uint16_t)map.cell(v) : (uint16_t)cells;
    }
    __syncthreads();
    // Each warp fills one contiguous run of cells, lanes interleaved (coalesced writes).  The cells
    // of a lane ascend, and so does k = #{borders in earlier cells}: one binary search for the
    // first cell, then k only ever advances -- by the few borders that fall between two cells of
    // the lane.  (A binary search per cell made the 4096-cell table of the 8-bit kernels cost
    // ~5 us per CTA, a tenth of the kernel.)
    int shared_cell = 0;
    const int per_warp = (cells + kWarps - 1) / kWarps;
    const int first = (threadIdx.x >> 5) * per_warp, last = min(cells, first + per_warp);
    int c = first + (threadIdx.x & 31);
    if (c < last) {
        int k = 0;  // first border whose cell is >= c
#pragma unroll
        for (int step = kSize >> 1; step >= 1; step >>= 1)
            if (s.bound_cell[k + step - 1] < c) k += step;
        for (; c < last; c += 32) {
            while (k < kSize && s.bound_cell[k] < c) ++k;
            shared_cell |= (k + 1 < kSize && s.bound_cell[k + 1] == c) ? 1 : 0;
            emit(c, k, k < kSize && s.bound_cell[k] == c);
        }
    }
    return __syncthreads_or(shared_cell) != 0;
}

// Two layouts behind one interface; which one serves (T, B) is decided by kOneGather below.
//
// fp32 -- two gathers: a byte per cell holding 4k (k at 7 and 8 bits, where 4k does not fit) and
// the borders themselves; code = k + (bounds[k] < x).  128 one-byte cells are 32 words, one per
// bank, and the 8 borders of a 3-bit table sit in 8 different banks: at 3 bits neither gather has
// a bank conflict.  The four k of a half are packed before the comparison results are added in, so
// a code costs one multiply-add and one predicated add.
//
// bf16 -- one gather: a 32-bit word per cell.  Borders of a bf16 table are bf16 values, so the low
// 16 bits of a border's float pattern are free.  The word is built so that, read as a float, it
// lies in [border, next bf16 above border) -- no bf16 input falls strictly inside, so
// (word < x) == (border < x) for every bf16 x -- while its low 16 bits carry k = #{borders in
// earlier cells}, replicated at every position a code can take inside a packed half (B <= 4: four
// copies, at bits 0, B, 2B, 3B; wider codes: two copies, the half is assembled from two pairs).
// Packing the k of element j is then ONE logic op, acc |= word & (mask << B j), and the code is
// completed by a predicated add of 1 << B j.  Cells without a border hold +inf or a NaN pattern
// (never below x).
template <typename T, int B> struct Bucketizer<T, B, false> {
    static constexpr bool kOneGather = sizeof(T) == 2;
    static constexpr int kSize = 1 << B;  // borders padded with +inf to a power of two
    static constexpr int kCells = kOneGather ? (B == 6 ? 2048 : 16 << B)      // 128 256 512 2048 2048 4096
                                            : B == 3 ? 128 : B == 4 ? 256 : B == 5 ? 512 : B == 8 ? 4096 : 2048;
    static constexpr int kShift = B <= 6 ? 2 : 0;   // byte entries hold k << kShift
    static constexpr int kCopies = B <= 4 ? 4 : 2;
    struct Scratch {
        float bounds[kSize];
        uint16_t bound_cell[kSize];
        alignas(4) uint8_t lut[kOneGather ? 4 : kCells];
        uint32_t word[kOneGather ? kCells : 1];
    };
    const T *bounds;
    int nbounds;
    CellMap map;
    bool crowded;  // some cell holds >= 2 borders: the whole block searches exactly

    static __device__ __forceinline__ uint32_t encode(float border, uint32_t k) {
        uint32_t low = k | (k << B);
        if (kCopies == 4) low |= low << (2 * B);
        const uint32_t bits = __float_as_uint(border);
        if ((bits << 1) == 0) return low;           // +-0: a denormal below every positive bf16
        // negative borders: the values of (border, next bf16 towards zero) have the next bf16's
        // high half; with nothing to carry the border itself will do
        if ((bits >> 31) && low != 0) return (bits - 0x10000u) | low;
        return bits | low;                          // (+inf: a NaN pattern unless k == 0)
    }
    // bf16 at 3 and 4 bits: a table of 32 words has one word per bank -- the gather is free of
    // bank conflicts (with 128 words ncu counted 2.6 wavefronts per gather, and the shared-memory
    // pipe is what bounds these kernels).  Most shipped 3-bit tables separate at 32 cells; if a
    // table does not, the full-size table is built instead.
    // The table is built with the smallest cell count, from kMinCells up in powers of two, that keeps any two
    // borders apart (choose_cells: one block-wide vote per candidate): 12 of the 13 shipped tables of 5-8 bits
    // separate at half of kCells or less, and a table half the size is built in half the time and gathered with
    // fewer bank conflicts (more lanes share a word) -- bf16 GELU 7 bits 79 -> 82 %, 8 bits 76 -> 80 % of the HBM
    // peak (profiles/r02_adaptive_cells.txt).  kCells is the size that every shipped table separates at.
    static constexpr int kMinCells = B >= 5 ? (FEWBIT_ADAPTIVE_CELLS ? 4 << B : kCells)
                                            : (FEWBIT_FEW_CELLS && kOneGather && B >= 3) ? 32 : kCells;
    __device__ __forceinline__ void prepare(Scratch &s) {
        map.fit(bounds, nbounds);
        auto emit = [&](int c, int k, bool has) {
            if constexpr (kOneGather)
                s.word[c] = encode(has ? s.bounds[k] : CUDART_INF_F, (uint32_t)k);
            else
                s.lut[c] = (uint8_t)(k << kShift);
        };
        const int cells = choose_cells(map, bounds, nbounds, kMinCells, kCells);
        crowded = build_cells<kSize>(map, cells, bounds, nbounds, s, emit);
    }
    __device__ __forceinline__ uint32_t half(const Scratch &s, float x0, float x1, float x2, float x3) const {
        uint32_t c0, c1, c2, c3;
        map.cells(x0, x1, c0, c1);
        map.cells(x2, x3, c2, c3);
        uint32_t acc;
        float b0, b1, b2, b3;
        if constexpr (kOneGather) {
            const uint32_t w0 = s.word[c0], w1 = s.word[c1], w2 = s.word[c2], w3 = s.word[c3];
            constexpr uint32_t m = (1u << B) - 1u;
            if constexpr (kCopies == 4) {
                acc = (w0 & m) | (w1 & (m << B)) | (w2 & (m << (2 * B))) | (w3 & (m << (3 * B)));
            } else {
                acc = ((w0 & m) | (w1 & (m << B))) + (((w2 & m) | (w3 & (m << B))) << (2 * B));
            }
            b0 = __uint_as_float(w0), b1 = __uint_as_float(w1), b2 = __uint_as_float(w2), b3 = __uint_as_float(w3);
        } else {
            const uint32_t e0 = s.lut[c0], e1 = s.lut[c1], e2 = s.lut[c2], e3 = s.lut[c3];
            const char *table = reinterpret_cast<const char *>(s.bounds);
            b0 = *reinterpret_cast<const float *>(table + (e0 << (2 - kShift)));
            b1 = *reinterpret_cast<const float *>(table + (e1 << (2 - kShift)));
            b2 = *reinterpret_cast<const float *>(table + (e2 << (2 - kShift)));
            b3 = *reinterpret_cast<const float *>(table + (e3 << (2 - kShift)));
            acc = pack4<B>(e0, e1, e2, e3) >> kShift;
        }
        add_if_less<1u>(acc, b0, x0);
        add_if_less<1u << B>(acc, b1, x1);
        add_if_less<1u << (2 * B)>(acc, b2, x2);
        add_if_less<1u << (3 * B)>(acc, b3, x3);
        return acc;
    }
    static constexpr bool kMayCrowd = true;
    template <bool kExact>
    __device__ __forceinline__ void lookup(const Scratch &s, const float (&x)[8], uint32_t (&out)[2]) const {
        if constexpr (kExact) {
            uint32_t code[8];
            const uint32_t sorted = (uint32_t)__cvta_generic_to_shared(s.bounds);
#pragma unroll
            for (int j = 0; j < 8; ++j) code[j] = search_exact<kSize>(sorted, x[j]);
            out[0] = pack4<B>(code[0], code[1], code[2], code[3]);
            out[1] = pack4<B>(code[4], code[5], code[6], code[7]);
        } else {
            out[0] = half(s, x[0], x[1], x[2], x[3]);
            out[1] = half(s, x[4], x[5], x[6], x[7]);
        }
    }
};

// =====================================================================================
// Continuous activations.  fp32: the expressions ATen's CUDA kernels use in fp32 opmath (that is
// what the reference's test compares against: functional/activations_test.py:88-89) with accurate
// functions -- libdevice, or `namespace accurate` below for expm1 and log1p (mish and softplus
// rearranged); every result within 3 ulp of ATen's, gelu bit for bit.
// bf16: inputs are widened to fp32, evaluated with the MUFU-based forms below (a bf16 kernel
// moves half the bytes per element, so libdevice-grade math would make it compute-bound at
// 26-50 % of HBM speed) and rounded once.  Reference lambdas: fewbit/cuda/codec.cu:517-653.
// =====================================================================================

// MUFU-based building blocks of the bf16 paths.  A bf16 result keeps 8 significant bits, so
// the ~2^-22 relative error of ex2/rcp/lg2.approx is invisible after the final rounding; what
// matters is the ALGEBRA: forms without cancellation, so the relative error stays ~1e-6
// everywhere (tests: <= 1 bf16 ulp from the fp32-exact result, and equal to it for > 99 %).
__device__ __forceinline__ float rcp_approx(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float ex2_approx(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float lg2_approx(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// Packed fp32 pairs.  sm_100 has FFMA2: one issue slot, two fp32 FMAs on an aligned register pair
// (same rounding as two FFMAs; |.| and negation fold into operand modifiers, equal halves into a
// scalar-broadcast operand).  The transcendental kernels are bound by instruction issue, not by
// the FMA pipe, so evaluating two elements per instruction is what buys time there.
struct Pair {
    unsigned long long bits;
};
__device__ __forceinline__ Pair pair(float lo, float hi) {
    Pair r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.bits) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpair(Pair p, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p.bits));
}

// Lane-generic arithmetic: every bf16 formula below is written once over L = float (one element:
// ragged kernel) or L = Pair (two elements per instruction: tile kernel).  Same operations in the
// same order, so both give the same bits.  Multiply-adds are spelled out (fma_) rather than left
// to contraction, for the same reason.
template <class L> __device__ __forceinline__ L lane(float c);
template <> __device__ __forceinline__ float lane<float>(float c) { return c; }
template <> __device__ __forceinline__ Pair lane<Pair>(float c) { return pair(c, c); }

__device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ Pair fma_(Pair a, Pair b, Pair c) {
    Pair r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.bits) : "l"(a.bits), "l"(b.bits), "l"(c.bits));
    return r;
}
__device__ __forceinline__ Pair mul_(Pair a, Pair b) {
    Pair r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.bits) : "l"(a.bits), "l"(b.bits));
    return r;
}
__device__ __forceinline__ Pair add_(Pair a, Pair b) {
    Pair r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.bits) : "l"(a.bits), "l"(b.bits));
    return r;
}
// Per-half operations of a Pair (MUFU, sign games, comparisons have no packed form).
#define FEWBIT_HALVES_1(NAME, EXPR)                         \
    __device__ __forceinline__ Pair NAME(Pair a) {          \
        float v, w;                                         \
        unpair(a, v, w);                                    \
        float lo, hi;                                       \
        { const float x = v; lo = (EXPR); }                 \
        { const float x = w; hi = (EXPR); }                 \
        return pair(lo, hi);                                \
    }                                                       \
    __device__ __forceinline__ float NAME(float x) { return (EXPR); }
FEWBIT_HALVES_1(rcp_, rcp_approx(x))
FEWBIT_HALVES_1(ex2_, ex2_approx(x))
FEWBIT_HALVES_1(lg2_, lg2_approx(x))
FEWBIT_HALVES_1(abs_, fabsf(x))
FEWBIT_HALVES_1(neg_, -x)
FEWBIT_HALVES_1(expf_, expf(x))                 // libdevice, 2 ulp
// 2^k from km = k + 1.5 * 2^23 (the "magic" rounding constant): k sits in the low mantissa bits
FEWBIT_HALVES_1(pow2_of_magic, __uint_as_float((__float_as_uint(x) << 23) + 0x3f800000u))
#undef FEWBIT_HALVES_1
__device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float copysign_(float mag, float sgn) { return copysignf(mag, sgn); }
// a < b ? t : f   and   a > b ? t : f
__device__ __forceinline__ float if_less(float a, float b, float t, float f) { return a < b ? t : f; }
__device__ __forceinline__ float if_greater(float a, float b, float t, float f) { return a > b ? t : f; }
__device__ __forceinline__ Pair min_(Pair a, Pair b) {
    float a0, a1, b0, b1;
    unpair(a, a0, a1), unpair(b, b0, b1);
    return pair(fminf(a0, b0), fminf(a1, b1));
}
__device__ __forceinline__ Pair copysign_(Pair mag, Pair sgn) {
    float m0, m1, s0, s1;
    unpair(mag, m0, m1), unpair(sgn, s0, s1);
    return pair(copysignf(m0, s0), copysignf(m1, s1));
}
__device__ __forceinline__ Pair if_less(Pair a, Pair b, Pair t, Pair f) {
    float a0, a1, b0, b1, t0, t1, f0, f1;
    unpair(a, a0, a1), unpair(b, b0, b1), unpair(t, t0, t1), unpair(f, f0, f1);
    return pair(a0 < b0 ? t0 : f0, a1 < b1 ? t1 : f1);
}
__device__ __forceinline__ Pair if_greater(Pair a, Pair b, Pair t, Pair f) {
    float a0, a1, b0, b1, t0, t1, f0, f1;
    unpair(a, a0, a1), unpair(b, b0, b1), unpair(t, t0, t1), unpair(f, f0, f1);
    return pair(a0 > b0 ? t0 : f0, a1 > b1 ? t1 : f1);
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
namespace fast {   // the bf16 formulas (see the accuracy note above); L = float or Pair
template <class L> __device__ __forceinline__ L exp(L x) { return ex2_(mul_(x, lane<L>(kLog2e))); }
template <class L> __device__ __forceinline__ L exp_neg(L x) { return ex2_(mul_(x, lane<L>(-kLog2e))); }
// log(1 + e^z): for z < -8 the sum 1 + e^z loses e^z's digits, and log1p(e^z) = e^z (1 - e^z/2 ..)
template <class L> __device__ __forceinline__ L log1p_exp(L z) {
    const L e = exp(z);
    return if_less(z, lane<L>(-8.0f), e, mul_(lg2_(add_(e, lane<L>(1.0f))), lane<L>(kLn2)));
}
// e^z - 1 without cancellation near zero (|z| < 2^-9: z + z^2/2).
template <class L> __device__ __forceinline__ L expm1(L z) {
    return if_less(abs_(z), lane<L>(0.001953125f), fma_(mul_(z, lane<L>(0.5f)), z, z),
                   add_(exp(z), lane<L>(-1.0f)));
}
// tanh(x) = 1 - 2 / (1 + e^{2x}) for either sign: e^{2x} = inf gives 1, e^{2x} = 0 gives -1, NaN stays
// NaN; no absolute value, no copysign.  Near zero the subtraction cancels (absolute error ~1e-7):
// below 2^-5 the result is x itself (tanh x = x (1 - x^2/3 ..), x^2/3 < 2^-11.5, far inside half a
// bf16 ulp), at 2^-5 the cancellation costs 3e-6 relative.
template <class L> __device__ __forceinline__ L tanh_big(L x) {
    const L e = ex2_(mul_(x, lane<L>(2.0f * kLog2e)));
    return fma_(rcp_(add_(e, lane<L>(1.0f))), lane<L>(-2.0f), lane<L>(1.0f));
}
template <class L> __device__ __forceinline__ L tanh(L x) {
    return if_less(abs_(x), lane<L>(0.03125f), x, tanh_big(x));
}
// x - tanh(x) cancels near zero: below 1/8 the series x^3 (1/3 - 2 x^2/15) (next term 17 x^7/315:
// 2e-5 relative at 1/8); above, the subtraction keeps 1e-7 / (x^3/3) < 2e-4 relative.
template <class L> __device__ __forceinline__ L tanhshrink(L x) {
    const L x2 = mul_(x, x);
    const L p = fma_(x2, lane<L>(-0.13333333333f), lane<L>(0.33333333333f));
    return if_less(abs_(x), lane<L>(0.125f), mul_(mul_(x, x2), p), add_(x, neg_(tanh_big(x))));
}
template <class L> __device__ __forceinline__ L sigmoid(L x) { return rcp_(add_(exp_neg(x), lane<L>(1.0f))); }
template <class L> __device__ __forceinline__ L silu(L x) { return mul_(x, sigmoid(x)); }
template <class L> __device__ __forceinline__ L softsign(L x) { return mul_(x, rcp_(add_(abs_(x), lane<L>(1.0f)))); }
template <class L> __device__ __forceinline__ L logsigmoid(L x) {
    return add_(min_(x, lane<L>(0.0f)), neg_(log1p_exp(neg_(abs_(x)))));
}
// tanh(log(1 + e)) = n / (n + 2) with n = e (e + 2), e = e^x (capped: 1 in fp32 beyond x = 20)
template <class L> __device__ __forceinline__ L mish(L x) {
    const L e = exp(min_(x, lane<L>(20.0f))), n = mul_(e, add_(e, lane<L>(2.0f)));
    return mul_(mul_(x, n), rcp_(add_(n, lane<L>(2.0f))));
}
template <class L> __device__ __forceinline__ L softplus(L x, float beta, float threshold, float inv_beta) {
    const L bx = mul_(x, lane<L>(beta));
    return if_greater(bx, lane<L>(threshold), x, mul_(log1p_exp(bx), lane<L>(inv_beta)));
}
template <class L> __device__ __forceinline__ L elu(L x, float pos, float neg, float in_scale) {
    return if_greater(x, lane<L>(0.0f), mul_(x, lane<L>(pos)),
                      mul_(expm1(mul_(x, lane<L>(in_scale))), lane<L>(neg)));
}
// GELU through one branch-free erfc (derivation at GeluFn).
template <class L> __device__ __forceinline__ L gelu(L x) {
    const L one = lane<L>(1.0f);
    const L s = rcp_(fma_(abs_(x), lane<L>(0.35355339059327376220f), one));
    L p = fma_(lane<L>(2.816799879e-01f), s, lane<L>(-8.819190860e-01f));
    p = fma_(p, s, lane<L>(5.309718251e-01f));
    p = fma_(p, s, lane<L>(4.433360100e-01f));
    p = fma_(p, s, lane<L>(1.451876998e+00f));
    p = fma_(p, s, lane<L>(-1.825967312e+00f));
    // erfc(t) = s * 2^(P(s) - log2(e) t^2),  log2(e) t^2 = (log2(e)/2) x^2
    const L e = ex2_(fma_(mul_(x, lane<L>(-0.72134752044448170368f)), x, p));
    // (x/2)(1 + erf(x/sqrt 2)) with erf = sign(x)(1 - erfc):  x/2 + |x/2| (1 - s e)
    const L half = mul_(x, lane<L>(0.5f));
    return fma_(abs_(half), fma_(neg_(s), e, one), half);
}
}  // namespace fast

namespace accurate {   // fp32-grade replacements of the two costly libdevice calls; L = float or Pair
// Coefficients and measured error bounds: tools/fit_fp32_math.py.
//
// e^z - 1 for z <= 0 (the only range ELU, CELU and SELU need; larger z gives garbage, NaN stays
// NaN): z = k ln2 + r, |r| <= ln2/2, expm1(r) = r + r^2 Q(r), and
// e^z - 1 = 2^k expm1(r) + (2^k - 1) in one fused multiply-add -- 2^k - 1 is exact.  0.86 ulp
// against 29 instructions of expm1f; 13 of the 16 operations here pair up.
template <class L> __device__ __forceinline__ L expm1_nonpos(L z) {
    z = if_less(z, lane<L>(-88.0f), lane<L>(-88.0f), z);                 // 2^k would leave the exponent range
    const L magic = lane<L>(12582912.0f);
    const L km = fma_(z, lane<L>(1.4426950408889634f), magic);           // k = rint(z log2 e), biased
    const L k = add_(km, lane<L>(-12582912.0f));
    L r = fma_(k, lane<L>(-0.693145751953125f), z);                      // Cody-Waite: ln2 = hi + lo
    r = fma_(k, lane<L>(-1.428606765330187e-06f), r);
    L q = fma_(lane<L>(1.990645978e-04f), r, lane<L>(1.394211431e-03f));
    q = fma_(q, r, lane<L>(8.333287202e-03f));
    q = fma_(q, r, lane<L>(4.166634008e-02f));
    q = fma_(q, r, lane<L>(1.666666716e-01f));
    q = fma_(q, r, lane<L>(5.000000000e-01f));
    const L p = fma_(mul_(r, r), q, r);
    const L t = pow2_of_magic(km);
    return fma_(t, p, add_(t, lane<L>(-1.0f)));
}
// log(1 + e) for e in (0, 1] (softplus and logsigmoid call it with e = exp(-|x|)):
// 2 atanh(s), s = e / (2 + e) <= 1/3, as 2 s + s^3 R(s^2); the quotient from MUFU.RCP, one Newton
// step and a residual correction (correctly rounded s).  1.1 ulp against ~30 instructions of log1pf.
template <class L> __device__ __forceinline__ L log1p_unit(L e) {
    const L d = add_(e, lane<L>(2.0f)), one = lane<L>(1.0f);
    const L y0 = rcp_(d);
    const L y = fma_(fma_(neg_(d), y0, one), y0, y0);
    L s = mul_(e, y);
    // residual e - s (2 + e) against the TRUE denominator (d itself is rounded): e - 2 s is exact
    s = fma_(fma_(neg_(s), e, fma_(lane<L>(-2.0f), s, e)), y, s);
    const L u = mul_(s, s);
    L q = fma_(lane<L>(2.093757242e-01f), u, lane<L>(1.742623448e-01f));
    q = fma_(q, u, lane<L>(2.226814181e-01f));
    q = fma_(q, u, lane<L>(2.857018113e-01f));
    q = fma_(q, u, lane<L>(4.000001252e-01f));
    q = fma_(q, u, lane<L>(6.666666865e-01f));
    return fma_(mul_(s, u), q, add_(s, s));
}
template <class L> __device__ __forceinline__ L elu(L x, float pos, float neg, float in_scale) {
    return if_greater(x, lane<L>(0.0f), mul_(x, lane<L>(pos)),
                      mul_(expm1_nonpos(mul_(x, lane<L>(in_scale))), lane<L>(neg)));
}
// a / d for a finite d with 2^-120 < |d| < 2^120: MUFU.RCP, one Newton step, and the quotient's
// residual folded back in -- the correctly rounded result in all but rare ties, six pairable
// operations against the ~9 (plus a slow-path check) of the IEEE division sequence.
template <class L> __device__ __forceinline__ L quotient(L a, L d) {
    const L y0 = rcp_(d);
    const L y = fma_(fma_(neg_(d), y0, lane<L>(1.0f)), y0, y0);
    const L q = mul_(a, y);
    return fma_(fma_(neg_(d), q, a), y, q);
}
// 1 + e^-x, with -x capped at 80 so that the denominator stays inside quotient()'s range; below
// x = -80 the results differ from the exact ones by less than 2e-33 (NaN stays NaN).
template <class L> __device__ __forceinline__ L one_plus_exp_neg(L x) {
    const L z = neg_(x);
    return add_(expf_(if_greater(z, lane<L>(80.0f), lane<L>(80.0f), z)), lane<L>(1.0f));
}
template <class L> __device__ __forceinline__ L sigmoid(L x) { return quotient(lane<L>(1.0f), one_plus_exp_neg(x)); }
// x / (1 + e^-x).  An infinite numerator would turn the residual step into inf - inf: +inf stays
// +inf and -inf gives NaN (-inf / inf), as ATen's and the reference's expression do.
template <class L> __device__ __forceinline__ L silu(L x) {
    const L special = if_greater(x, lane<L>(0.0f), x, lane<L>(CUDART_NAN_F));
    return if_less(abs_(x), lane<L>(CUDART_INF_F), quotient(x, one_plus_exp_neg(x)), special);
}
// x n / (n + 2), n = e (e + 2), e = e^min(x, 20) <= 4.9e8: the denominator is at most 2.4e17
template <class L> __device__ __forceinline__ L mish(L x) {
    const L e = expf_(min_(x, lane<L>(20.0f))), n = mul_(e, add_(e, lane<L>(2.0f)));
    return mul_(x, quotient(n, add_(n, lane<L>(2.0f))));
}
// min(0, x) - log1p(exp(-|x|)): ATen's own expression
template <class L> __device__ __forceinline__ L logsigmoid(L x) {
    return add_(min_(x, lane<L>(0.0f)), neg_(log1p_unit(expf_(neg_(abs_(x))))));
}
// log1p(exp(z)) = max(z, 0) + log1p(exp(-|z|)), z = beta x: no overflow, and the logarithm's
// argument stays in (1, 2]
template <class L> __device__ __forceinline__ L softplus_of_scaled(L z) {
    return add_(if_greater(z, lane<L>(0.0f), z, lane<L>(0.0f)), log1p_unit(expf_(neg_(abs_(z)))));
}
}  // namespace accurate

// A functor with a bf16 formula exposes it for one element and for a pair.
// PAIRED says whether the tile kernel uses the pair form: measured per function on B200
// (profiles/r01_function_sweep_3bit.md) -- it pays where the formula is mostly multiply-adds
// (gelu +9 %, tanhshrink +9 %, tanh, mish, elu family) and not where it is mostly MUFU and selects.
#define FEWBIT_BF16_FORMULA(PAIRED, CALL1, CALL2)                                              \
    static constexpr bool kPaired = PAIRED;                                                    \
    __device__ __forceinline__ float bf16_value(float x) const { return CALL1; }              \
    __device__ __forceinline__ void bf16_pair(float &x0, float &x1) const {                    \
        const Pair x = pair(x0, x1);                                                           \
        unpair(CALL2, x0, x1);                                                                 \
    }

// The same for an fp32 formula built from the accurate:: helpers (always used in pair form by the
// tile kernel: these formulas are multiply-add chains).
#define FEWBIT_F32_FORMULA(CALL1, CALL2)                                                       \
    static constexpr bool kPaired32 = true;                                                    \
    __device__ __forceinline__ float f32_value(float x) const { return CALL1; }               \
    __device__ __forceinline__ void f32_pair(float &x0, float &x1) const {                     \
        const Pair x = pair(x0, x1);                                                           \
        unpair(CALL2, x0, x1);                                                                 \
    }

struct EluFamily {  // celu / elu / selu:  x > 0 ? x*pos : expm1(x*in_scale)*neg
    float pos, neg, in_scale;
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return bf16_value(x);
        return f32_value(x);
    }
    FEWBIT_BF16_FORMULA(true, fast::elu<float>(x, pos, neg, in_scale), fast::elu<Pair>(x, pos, neg, in_scale))
    FEWBIT_F32_FORMULA(accurate::elu<float>(x, pos, neg, in_scale), accurate::elu<Pair>(x, pos, neg, in_scale))
};
struct CeluFn : EluFamily {  // codec.cu:517-526
    __host__ CeluFn(double alpha, double) : EluFamily{1.0f, (float)alpha, (float)(1.0 / alpha)} {}
};
struct EluFn : EluFamily {  // codec.cu:528-537
    __host__ EluFn(double alpha, double) : EluFamily{1.0f, (float)alpha, 1.0f} {}
};
struct SeluFn : EluFamily {  // codec.cu:588-600
    __host__ SeluFn(double, double)
        : EluFamily{(float)1.0507009873554804934193349852946,
                    (float)1.6732632423543772848170429916717 *
                        (float)1.0507009873554804934193349852946,
                    1.0f} {}
};
// GELU.  ATen evaluates  (0.5 x) * (1 + erf(x / sqrt 2))  with libdevice erff, a two-branch
// routine that costs ~9 ALU-pipe selects per element -- affordable in fp32, where the kernel is
// close to the HBM bound anyway, but not in bf16 (half the bytes per element).  For bf16 erf
// comes from one branch-free evaluation of erfc:
//     erfc(t) = s * 2^(P(s) - log2(e) t^2),   s = 1 / (1 + t/2),   t = |x| / sqrt 2
// with a degree-5 P fitted by tools/fit_gelu.py (0.1 % of results differ from the fp32-exact
// bf16 rounding, by one bf16 ulp): 11 FMA-pipe ops + 2 MUFU, no ALU-pipe op.
struct GeluFn {  // codec.cu:539-544 (x * normcdf(x))
    __host__ GeluFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 4) {
            // fp32: ATen's expression with libdevice erff, bit for bit F.gelu.  (A more accurate
            // erf lands on a different point of the 2^-24 grid of 1 + erf than erff does in the
            // negative tail, and the reference's own test -- L2 distance to F.gelu below 1e-6
            // on linspace(-5, 5, 101) -- then fails at 1.1e-6; see DESIGN.md.)
            return (x * 0.5f) * (1.0f + erff(x * 0.70710678118654752440f));
        } else {
            return bf16_value(x);   // rounded to 8 bits: the short branch-free erfc is enough
        }
    }
    FEWBIT_BF16_FORMULA(true, fast::gelu<float>(x), fast::gelu<Pair>(x))
};
struct HardswishFn {  // codec.cu:546-564
    __host__ HardswishFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        // bf16: x * sat(x/6 + 1/2) -- the clamp rides on the multiply-add (FFMA.SAT, FMA pipe);
        // min/max cost two ALU-pipe ops per element, and the ALU pipe is what bounds this kernel
        // (ncu: 68 % busy).  Same value as ATen's expression up to the rounding of x/6 + 1/2, far
        // below bf16 resolution; NaN and infinities behave alike (-inf * 0 = NaN in both).
        if constexpr (sizeof(T) == 2) return x * __saturatef(fmaf(x, 1.0f / 6.0f, 0.5f));
        return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
    }
};
struct LogSigmoidFn {  // codec.cu:566-576
    __host__ LogSigmoidFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return bf16_value(x);
        return f32_value(x);
    }
    FEWBIT_BF16_FORMULA(true, fast::logsigmoid<float>(x), fast::logsigmoid<Pair>(x))
    FEWBIT_F32_FORMULA(accurate::logsigmoid<float>(x), accurate::logsigmoid<Pair>(x))
};
struct MishFn {  // codec.cu:578-586
    // tanh(log(1 + e)) = ((1+e)^2 - 1) / ((1+e)^2 + 1) = n / (n + 2) with n = e (e + 2), e = e^x:
    // one exponential and one division instead of exp, log1p and tanh (57 -> ~25 instructions),
    // all terms positive (no cancellation).  x is capped at 20 in the exponential, where
    // n / (n + 2) is already 1 in fp32.
    __host__ MishFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return bf16_value(x);
        return f32_value(x);
    }
    FEWBIT_BF16_FORMULA(true, fast::mish<float>(x), fast::mish<Pair>(x))
    FEWBIT_F32_FORMULA(accurate::mish<float>(x), accurate::mish<Pair>(x))
};
struct SigmoidFn {  // codec.cu:602-607
    __host__ SigmoidFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return bf16_value(x);
        return f32_value(x);
    }
    FEWBIT_BF16_FORMULA(true, fast::sigmoid<float>(x), fast::sigmoid<Pair>(x))
    FEWBIT_F32_FORMULA(accurate::sigmoid<float>(x), accurate::sigmoid<Pair>(x))
};
struct SiluFn {  // codec.cu:609-614
    __host__ SiluFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return bf16_value(x);
        return f32_value(x);
    }
    FEWBIT_BF16_FORMULA(true, fast::silu<float>(x), fast::silu<Pair>(x))
    FEWBIT_F32_FORMULA(accurate::silu<float>(x), accurate::silu<Pair>(x))
};
struct SoftplusFn {  // codec.cu:616-632
    float beta, threshold, inv_beta;
    bool unit_beta;
    __host__ SoftplusFn(double b, double t)
        : beta((float)b), threshold((float)t), inv_beta((float)(1.0 / b)), unit_beta((float)b == 1.0f) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return bf16_value(x);
        return f32_value(x);
    }
    // ATen: x * beta > threshold ? x : log1p(exp(x * beta)) / beta  (x / 1 is exact: skipped)
    __device__ __forceinline__ float f32_value(float x) const {
        const float bx = x * beta, soft = accurate::softplus_of_scaled<float>(bx);
        return bx > threshold ? x : (unit_beta ? soft : soft / beta);
    }
    static constexpr bool kPaired32 = true;
    __device__ __forceinline__ void f32_pair(float &x0, float &x1) const {
        const Pair bx = mul_(pair(x0, x1), lane<Pair>(beta));
        float b0, b1, s0, s1;
        unpair(bx, b0, b1);
        unpair(accurate::softplus_of_scaled<Pair>(bx), s0, s1);
        if (!unit_beta) s0 = s0 / beta, s1 = s1 / beta;
        x0 = b0 > threshold ? x0 : s0;
        x1 = b1 > threshold ? x1 : s1;
    }
    FEWBIT_BF16_FORMULA(false, fast::softplus<float>(x, beta, threshold, inv_beta),
                        fast::softplus<Pair>(x, beta, threshold, inv_beta))
};
struct SoftsignFn {  // codec.cu:634-639
    __host__ SoftsignFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return bf16_value(x);
        return x / (1.0f + fabsf(x));
    }
    FEWBIT_BF16_FORMULA(false, fast::softsign<float>(x), fast::softsign<Pair>(x))
};
struct TanhFn {  // codec.cu:641-646
    __host__ TanhFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return bf16_value(x);
        return tanhf(x);
    }
    FEWBIT_BF16_FORMULA(true, fast::tanh<float>(x), fast::tanh<Pair>(x))
};
struct TanhshrinkFn {  // codec.cu:648-653
    __host__ TanhshrinkFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return bf16_value(x);
        return x - tanhf(x);
    }
    FEWBIT_BF16_FORMULA(true, fast::tanhshrink<float>(x), fast::tanhshrink<Pair>(x))
};

// Forward op of a continuous activation: y = fn(x), code = bucket(x), eight values at a time.
#ifndef FEWBIT_PAIRS
#define FEWBIT_PAIRS 1
#endif
// Functions whose bf16 formula is so short that streaming the input (tile.cuh, ForwardStream) costs
// more in bookkeeping than the exposed load latency it removes, up to 4 bits (measured per function,
// profiles/r02_stream_per_function.txt: hardswish 87 % without against 80 % with, sigmoid 85 / 80,
// softsign 86 / 80; from 5 bits on streaming wins or ties everywhere).
template <class Fn> struct is_light_bf16 : std::false_type {};
template <class Fn, typename = void> struct has_pairs : std::false_type {};
template <class Fn> struct has_pairs<Fn, std::enable_if_t<Fn::kPaired>> : std::true_type {};
template <class Fn, typename = void> struct has_pairs32 : std::false_type {};
template <class Fn> struct has_pairs32<Fn, std::enable_if_t<Fn::kPaired32>> : std::true_type {};

// Tile shape of the fp32 forward kernel per function: (subtiles per warp tile, resident CTAs the
// register budget is set for).  Measured on B200 over {1,4}, {2,4}, {2,3}, {4,3} for every function
// (3-bit, 128x128x3072; profiles/r01_function_sweep_3bit.md): short formulas want few loads in flight
// per warp and four CTAs, libdevice-heavy ones the registers of three CTAs.
#ifndef FEWBIT_STREAM_F32
#define FEWBIT_STREAM_F32 1   // 0: no fp32 kernel streams its input; 1: those marked below; 2: all
#endif
template <class Fn> struct TileHintF32 {
    static constexpr int kSubtiles = 4, kMinBlocks = 3;               // gelu, logsigmoid
    static constexpr bool kStream = FEWBIT_STREAM_F32 >= 1;
};
#define FEWBIT_TILE_HINT_F32(FN, U, MINB, STREAM)                         \
    template <> struct TileHintF32<FN> {                                  \
        static constexpr int kSubtiles = U, kMinBlocks = MINB;            \
        static constexpr bool kStream = FEWBIT_STREAM_F32 >= 2 || (FEWBIT_STREAM_F32 == 1 && STREAM); \
    };
FEWBIT_TILE_HINT_F32(CeluFn, 1, 4, 0)
FEWBIT_TILE_HINT_F32(EluFn, 1, 4, 0)
FEWBIT_TILE_HINT_F32(SeluFn, 1, 4, 0)
FEWBIT_TILE_HINT_F32(HardswishFn, 1, 4, 0)
FEWBIT_TILE_HINT_F32(SoftsignFn, 2, 3, 0)
FEWBIT_TILE_HINT_F32(TanhFn, 2, 3, 0)
FEWBIT_TILE_HINT_F32(TanhshrinkFn, 2, 3, 0)
FEWBIT_TILE_HINT_F32(SigmoidFn, 2, 3, 0)
FEWBIT_TILE_HINT_F32(MishFn, 2, 3, 0)
FEWBIT_TILE_HINT_F32(SiluFn, 2, 3, 0)
FEWBIT_TILE_HINT_F32(SoftplusFn, 2, 3, 1)
#undef FEWBIT_TILE_HINT_F32

template <> struct is_light_bf16<HardswishFn> : std::true_type {};
template <> struct is_light_bf16<SigmoidFn> : std::true_type {};
template <> struct is_light_bf16<SoftsignFn> : std::true_type {};

template <class Fn, typename T, int B> struct QuantizeOp {
    static constexpr int kBits = B;
    static constexpr int kSubtilesF32 = TileHintF32<Fn>::kSubtiles, kMinBlocksF32 = TileHintF32<Fn>::kMinBlocks;
    static constexpr bool kStreamF32 = TileHintF32<Fn>::kStream && B >= 3;   // 1-2 bits: measured slower (89 -> 81 %)
    static constexpr bool kStreamInput = !(sizeof(T) == 2 && is_light_bf16<Fn>::value && B <= 4);
    static constexpr bool kHeavy = true;  // transcendental math: see TileConfig in launch.cuh
    using Scratch = typename Bucketizer<T, B>::Scratch;
    Fn fn;
    Bucketizer<T, B> bucket;
    __device__ __forceinline__ void prepare(Scratch &s) { bucket.prepare(s); }
    __device__ __forceinline__ bool exact() const {
        if constexpr (Bucketizer<T, B>::kMayCrowd) return bucket.crowded;
        return false;
    }
    template <bool kExact>
    __device__ __forceinline__ void apply(const Scratch &s, float (&v)[8], uint32_t (&half)[2]) const {
        bucket.template lookup<kExact>(s, v, half);
        if constexpr (sizeof(T) == 2 && has_pairs<Fn>::value && FEWBIT_PAIRS) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) fn.bf16_pair(v[j], v[j + 1]);
        } else if constexpr (sizeof(T) == 4 && has_pairs32<Fn>::value && FEWBIT_PAIRS) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) fn.f32_pair(v[j], v[j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fn.template eval<T>(v[j]);
        }
    }
};

// Forward op of the custom-table operator `stepwise` (reference schema fewbit/fewbit.cc:37, module
// fewbit/modules/activations.py:97-134; the reference declares it and ships no kernel).  The
// table IS the function: levels are the slopes of a continuous piecewise-linear activation with
// kinks at the borders, anchored at F(anchor) = 0,
//     F(x) = levels[k] * (x - bounds[k-1]) + F(bounds[k-1])   for bounds[k-1] < x <= bounds[k],
// so that the backward pass -- levels[code] * grad, the shared kernel -- is its exact derivative.
// y = fma(levels[code], x, intercept[code]) with the intercepts built per block from a prefix sum
// (in double) over the pieces.
template <typename T, int B> struct IntegralOp {
    static constexpr int kBits = B;
    static constexpr bool kHeavy = true;
    static constexpr bool kStreamInput = false;   // its tables leave no room for the input ring at 8 bits
    static constexpr int kLevels = 1 << B;
    struct Scratch {
        typename Bucketizer<T, B>::Scratch bucket;
        float slope[kLevels], intercept[kLevels];
        double reach[kLevels];      // reach[k] = F~(bounds[k]) with F~(bounds[0]) = 0
    };
    Bucketizer<T, B> bucket;
    const T *levels;
    int nlevels;
    float anchor;

    __device__ __forceinline__ void prepare(Scratch &s) {
        const int nb = bucket.nbounds, t = threadIdx.x;
        // piece k (1 <= k < nb) spans (bounds[k-1], bounds[k]]: inclusive scan of its rise
        for (int k = t; k < kLevels; k += blockDim.x) {
            s.slope[k] = k < nlevels ? to_float<T>(levels[k]) : 0.0f;
            s.reach[k] = (k >= 1 && k < nb)
                             ? (double)to_float<T>(levels[k]) *
                                   ((double)to_float<T>(bucket.bounds[k]) - (double)to_float<T>(bucket.bounds[k - 1]))
                             : 0.0;
        }
        __syncthreads();
        for (int step = 1; step < kLevels; step <<= 1) {
            double add[(kLevels + kThreads - 1) / kThreads];
#pragma unroll
            for (int i = 0; i < (kLevels + kThreads - 1) / kThreads; ++i) {
                const int k = t + i * kThreads;
                add[i] = (k < kLevels && k >= step) ? s.reach[k - step] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < (kLevels + kThreads - 1) / kThreads; ++i) {
                const int k = t + i * kThreads;
                if (k < kLevels) s.reach[k] += add[i];
            }
            __syncthreads();
        }
        // F~ at the anchor: its piece is found by the reference's own search (first border not below it)
        int piece = 0;
        while (piece < nb && to_float<T>(bucket.bounds[piece]) < anchor) ++piece;
        double at_anchor = 0.0;
        if (nb > 0) {
            const int left = piece == 0 ? 0 : piece - 1;   // F~ is known at bounds[left]
            at_anchor = s.reach[left] + (double)s.slope[piece] * ((double)anchor - (double)to_float<T>(bucket.bounds[left]));
        } else {
            at_anchor = (double)s.slope[0] * (double)anchor;
        }
        for (int k = t; k < kLevels; k += blockDim.x) {
            double c;
            if (nb == 0) {
                c = 0.0;
            } else {
                const int left = k == 0 ? 0 : (k - 1 < nb ? k - 1 : nb - 1);
                c = s.reach[left] - (double)s.slope[k] * (double)to_float<T>(bucket.bounds[left]);
            }
            s.intercept[k] = (float)(c - at_anchor);
        }
        bucket.prepare(s.bucket);      // ends with a block-wide barrier
    }
    __device__ __forceinline__ bool exact() const {
        if constexpr (Bucketizer<T, B>::kMayCrowd) return bucket.crowded;
        return false;
    }
    template <bool kExact>
    __device__ __forceinline__ void apply(const Scratch &s, float (&v)[8], uint32_t (&half)[2]) const {
        bucket.template lookup<kExact>(s.bucket, v, half);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t code = (half[j >> 2] >> (B * (j & 3))) & (kLevels - 1);
            v[j] = fmaf(s.slope[code], v[j], s.intercept[code]);
        }
    }
};

// Backward op of every continuous activation: factor = levels[code]
// (StepwiseBackwardKernel, fewbit/cuda/codec.cu:655-670).
template <typename T, int B> struct LevelsOp {
    static constexpr int kBits = B;
    static constexpr bool kHeavy = false;
    struct Scratch {
        float levels[1 << B];
    };
    const T *levels;
    int nlevels;
    const float *table;
    __device__ __forceinline__ void prepare(Scratch &s) {
        for (int i = threadIdx.x; i < (1 << B); i += blockDim.x)
            s.levels[i] = i < nlevels ? to_float<T>(levels[i]) : 0.0f;
        __syncthreads();
        table = s.levels;
    }
    __device__ __forceinline__ float factor(uint32_t code) const { return table[code]; }
};

// =====================================================================================
// Piecewise (1-bit) family.  Branch order follows the reference kernels so that NaN takes
// the same branch; values follow torch.nn.functional (reference test:
// functional/activations_test.py:17-68).  fewbit/cuda/codec.cu:298-487.
// =====================================================================================

struct HardshrinkFn {  // codec.cu:298-311
    float lambd;
    __host__ HardshrinkFn(double l, double) : lambd((float)l) {}
    __device__ __forceinline__ float operator()(float x, uint32_t &m) const {
        m = (x < -lambd || x > lambd) ? 1u : 0u;
        return m ? x : 0.0f;
    }
};
struct HardsigmoidFn {  // codec.cu:316-331
    __host__ HardsigmoidFn(double, double) {}
    __device__ __forceinline__ float operator()(float x, uint32_t &m) const {
        const bool low = x <= -3.0f, high = x >= 3.0f;
        m = (low || high) ? 0u : 1u;
        const float mid = (x + 3.0f) * (1.0f / 6.0f);
        return low ? 0.0f : (high ? 1.0f : mid);
    }
};
struct HardtanhFn {  // codec.cu:354-370
    float lo, hi;
    __host__ HardtanhFn(double a, double b) : lo((float)a), hi((float)b) {}
    __device__ __forceinline__ float operator()(float x, uint32_t &m) const {
        const bool low = x <= lo, high = x >= hi;
        m = (low || high) ? 0u : 1u;
        return low ? lo : (high ? hi : x);
    }
};
struct LeakyReluFn {  // codec.cu:375-389 -- the mask marks the NEGATIVE side
    float slope;
    __host__ LeakyReluFn(double s, double) : slope((float)s) {}
    __device__ __forceinline__ float operator()(float x, uint32_t &m) const {
        const bool pos = x >= 0.0f;
        m = pos ? 0u : 1u;
        return pos ? x : slope * x;
    }
};
struct ReluFn {  // codec.cu:412-425
    __host__ ReluFn(double, double) {}
    __device__ __forceinline__ float operator()(float x, uint32_t &m) const {
        const bool off = x <= 0.0f;
        m = off ? 0u : 1u;
        return off ? 0.0f : x;
    }
};
struct Relu6Fn {  // codec.cu:430-445; saturates at 6.0
    __host__ Relu6Fn(double, double) {}
    __device__ __forceinline__ float operator()(float x, uint32_t &m) const {
        const bool low = x <= 0.0f, high = x >= 6.0f;
        m = (low || high) ? 0u : 1u;
        return low ? 0.0f : (high ? 6.0f : x);
    }
};
struct SoftshrinkFn {  // codec.cu:450-465
    float lambd;
    __host__ SoftshrinkFn(double l, double) : lambd((float)l) {}
    __device__ __forceinline__ float operator()(float x, uint32_t &m) const {
        const bool neg = x < -lambd, pos = x > lambd;
        m = (neg || pos) ? 1u : 0u;
        // both shifts up front and two selects: left as nested ternaries the bf16 instantiation
        // compiled to divergent branches (59 % of HBM speed instead of 85 %)
        const float up = x + lambd, down = x - lambd;
        const float r = pos ? down : 0.0f;
        return neg ? up : r;
    }
};
struct ThresholdFn {  // codec.cu:470-484
    float threshold, value;
    __host__ ThresholdFn(double t, double v) : threshold((float)t), value((float)v) {}
    __device__ __forceinline__ float operator()(float x, uint32_t &m) const {
        const bool off = x <= threshold;
        m = off ? 0u : 1u;
        return off ? value : x;
    }
};

template <class Fn> struct MaskOp {
    static constexpr int kBits = 1;
    static constexpr bool kHeavy = false;
    using Scratch = NoScratch;
    Fn fn;
    __device__ __forceinline__ void prepare(Scratch &) {}
    __device__ __forceinline__ bool exact() const { return false; }
    template <bool>
    __device__ __forceinline__ void apply(const Scratch &, float (&v)[8], uint32_t (&half)[2]) const {
        uint32_t code[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fn(v[j], code[j]);
        half[0] = pack4<1>(code[0], code[1], code[2], code[3]);
        half[1] = pack4<1>(code[4], code[5], code[6], code[7]);
    }
};

// Backward of the 1-bit family: factor = mask ? on : off  (codec.cu:271-296: idx * g;
// hardsigmoid :333-345: 1/6 | 0; leaky_relu :391-402: slope | 1).
struct MaskFactorOp {
    static constexpr int kBits = 1;
    static constexpr bool kHeavy = false;
    using Scratch = NoScratch;
    float on, off;
    __device__ __forceinline__ void prepare(Scratch &) {}
    __device__ __forceinline__ float factor(uint32_t code) const { return code ? on : off; }
};

}  // namespace fewbit
