// fewbit_b200 -- per-element operators plugged into the warp-tile kernels of tile.cuh.
#pragma once

#include <math_constants.h>

#include <type_traits>

#include "tile.cuh"

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

    __device__ __forceinline__ void lookup(const float (&x)[8], uint32_t (&code)[8]) const {
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
    }
};

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

template <typename T, int B> struct Bucketizer<T, B, false> {
    static constexpr int kSize = 1 << B;  // bounds padded with +inf to a power of two
    // Cells: enough to separate the borders of every shipped table, few enough that the
    // gather stays (nearly) bank-conflict free: 128 one-byte entries = 32 words = 1 per bank.
    static constexpr int kCells = B == 3 ? 128 : B == 4 ? 256 : B == 5 ? 512 : 2048;
    struct Scratch {
        float bounds[kSize];
        uint16_t bound_cell[kSize];
        alignas(4) uint8_t lut[kCells];
    };
    const T *bounds;
    int nbounds;
    float scale, offset;
    float lut_base;       // shared-space byte address of lut[], as the bits of a denormal float
    uint32_t table_base;  // shared-space byte address of bounds[]
    const float *table;
    bool crowded;         // some cell holds >= 2 borders: the whole block searches exactly

    // Shared-memory address of x's LUT entry, computed entirely on the FMA pipe: sat() clamps
    // (and sends NaN to 0); the second FMA works in the denormal range, where the bit pattern
    // of a float *is* its value in units of 2^-149, so  bits = round(t * (C-1)) + base.
    __device__ __forceinline__ uint32_t entry_address(float x) const {
        const float t = __saturatef(fmaf(x, scale, offset));
        return __float_as_uint(fmaf(t, __uint_as_float((uint32_t)(kCells - 1)), lut_base));
    }

    __device__ __forceinline__ void prepare(Scratch &s) {
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
        const uint32_t lut_address = (uint32_t)__cvta_generic_to_shared(s.lut);
        lut_base = __uint_as_float(lut_address);
        table_base = (uint32_t)__cvta_generic_to_shared(s.bounds);
        table = s.bounds;
        for (int i = threadIdx.x; i < kSize; i += blockDim.x) {
            const float v = i < nbounds ? to_float<T>(bounds[i]) : CUDART_INF_F;
            s.bounds[i] = v;
            s.bound_cell[i] = i < nbounds ? (uint16_t)(entry_address(v) - lut_address) : (uint16_t)kCells;
        }
        __syncthreads();
        int shared_cell = 0;
        for (int c = threadIdx.x; c < kCells; c += blockDim.x) {
            int k = 0;  // first bound whose cell is >= c
#pragma unroll
            for (int step = kSize >> 1; step >= 1; step >>= 1)
                if (s.bound_cell[k + step - 1] < c) k += step;
            shared_cell |= (k + 1 < kSize && s.bound_cell[k + 1] == c) ? 1 : 0;
            s.lut[c] = (uint8_t)k;
        }
        crowded = __syncthreads_or(shared_cell) != 0;
    }

    // Exact branch-free binary search; out of line on purpose (rare, keeps the hot loop small).
    static __device__ __noinline__ uint32_t exact(const float *sorted, float x) {
        int k = 0;
#pragma unroll
        for (int step = kSize >> 1; step >= 1; step >>= 1)
            if (sorted[k + step - 1] < x) k += step;
        return (uint32_t)k;
    }

    __device__ __forceinline__ void lookup(const float (&x)[8], uint32_t (&code)[8]) const {
        if (crowded) {  // block-uniform
#pragma unroll
            for (int j = 0; j < 8; ++j) code[j] = exact(table, x[j]);
            return;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) code[j] = lds_u8(entry_address(x[j]));
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (lds_f32(table_base + 4 * code[j]) < x[j]) code[j] += 1;
    }
};

// =====================================================================================
// Continuous activations.  fp32: the expressions ATen's CUDA kernels use in fp32 opmath (that is
// what the reference's test compares against: functional/activations_test.py:88-89) with the
// accurate libdevice functions (mish and softplus/beta=1 rearranged, still within 4 ulp).
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
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float fast_exp(float x) { return ex2_approx(x * kLog2e); }
// log(1 + e^z): for z < -8 the sum 1 + e^z loses e^z's digits, and log1p(e^z) = e^z (1 - e^z/2 ..)
__device__ __forceinline__ float fast_log1p_exp(float z) {
    const float e = fast_exp(z);
    return z < -8.0f ? e : kLn2 * lg2_approx(1.0f + e);
}
// e^z - 1 without cancellation near zero (|z| < 2^-9: z + z^2/2).
__device__ __forceinline__ float fast_expm1(float z) {
    return fabsf(z) < 0.001953125f ? fmaf(0.5f * z, z, z) : fast_exp(z) - 1.0f;
}
// tanh: odd polynomial below 1/4 (relative error 3e-7), 1 - 2/(1 + e^{2|x|}) above.
__device__ __forceinline__ float fast_tanh(float x) {
    const float a = fabsf(x), x2 = x * x;
    const float small = x * fmaf(x2, fmaf(x2, fmaf(x2, -0.05396825397f, 0.13333333333f), -0.33333333333f), 1.0f);
    const float big = copysignf(fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(a * (2.0f * kLog2e))), 1.0f), x);
    return a < 0.25f ? small : big;
}

struct EluFamily {  // celu / elu / selu:  x > 0 ? x*pos : expm1(x*in_scale)*neg
    float pos, neg, in_scale;
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return x > 0.0f ? x * pos : fast_expm1(x * in_scale) * neg;
        return x > 0.0f ? x * pos : expm1f(x * in_scale) * neg;
    }
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
            // bf16: the result is rounded to 8 bits, so erfc may come from the short branch-free
            // evaluation above.  s = 1 / (1 + t/2) with t = |x| / sqrt 2.
            const float s = rcp_approx(fmaf(fabsf(x), 0.35355339059327376220f, 1.0f));
            float p = fmaf(2.816799879e-01f, s, -8.819190860e-01f);
            p = fmaf(p, s, 5.309718251e-01f);
            p = fmaf(p, s, 4.433360100e-01f);
            p = fmaf(p, s, 1.451876998e+00f);
            p = fmaf(p, s, -1.825967312e+00f);
            // erfc(t) = s * 2^(P(s) - log2(e) t^2),  log2(e) t^2 = (log2(e)/2) x^2
            const float erfc_t = s * ex2_approx(fmaf(x * -0.72134752044448170368f, x, p));
            // (x/2)(1 + erf(x/sqrt 2)) with erf = sign(x)(1 - erfc):  x/2 + |x/2| (1 - erfc)
            const float half = x * 0.5f;
            return fmaf(fabsf(half), 1.0f - erfc_t, half);
        }
    }
};
struct HardswishFn {  // codec.cu:546-564
    __host__ HardswishFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
    }
};
struct LogSigmoidFn {  // codec.cu:566-576
    __host__ LogSigmoidFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return fminf(0.0f, x) - fast_log1p_exp(-fabsf(x));
        return fminf(0.0f, x) - log1pf(expf(-fabsf(x)));
    }
};
struct MishFn {  // codec.cu:578-586
    // tanh(log(1 + e)) = ((1+e)^2 - 1) / ((1+e)^2 + 1) = n / (n + 2) with n = e (e + 2), e = e^x:
    // one exponential and one division instead of exp, log1p and tanh (57 -> ~25 instructions),
    // all terms positive (no cancellation).  x is capped at 20 in the exponential, where
    // n / (n + 2) is already 1 in fp32.
    __host__ MishFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) {
            const float e = fast_exp(fminf(x, 20.0f)), n = e * (e + 2.0f);
            return x * n * rcp_approx(n + 2.0f);
        }
        const float e = expf(fminf(x, 20.0f)), n = e * (e + 2.0f);
        return x * (n / (n + 2.0f));
    }
};
struct SigmoidFn {  // codec.cu:602-607
    __host__ SigmoidFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return rcp_approx(1.0f + fast_exp(-x));
        return 1.0f / (1.0f + expf(-x));
    }
};
struct SiluFn {  // codec.cu:609-614
    __host__ SiluFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return x * rcp_approx(1.0f + fast_exp(-x));
        return x / (1.0f + expf(-x));
    }
};
struct SoftplusFn {  // codec.cu:616-632
    float beta, threshold, inv_beta;
    bool unit_beta;
    __host__ SoftplusFn(double b, double t)
        : beta((float)b), threshold((float)t), inv_beta((float)(1.0 / b)), unit_beta((float)b == 1.0f) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        const float bx = x * beta;
        if constexpr (sizeof(T) == 2) return bx > threshold ? x : fast_log1p_exp(bx) * inv_beta;
        const float soft = log1pf(expf(bx));
        return bx > threshold ? x : (unit_beta ? soft : soft / beta);   // x / 1 is exact: skip it
    }
};
struct SoftsignFn {  // codec.cu:634-639
    __host__ SoftsignFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return x * rcp_approx(1.0f + fabsf(x));
        return x / (1.0f + fabsf(x));
    }
};
struct TanhFn {  // codec.cu:641-646
    __host__ TanhFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) return fast_tanh(x);
        return tanhf(x);
    }
};
struct TanhshrinkFn {  // codec.cu:648-653
    __host__ TanhshrinkFn(double, double) {}
    template <typename T> __device__ __forceinline__ float eval(float x) const {
        if constexpr (sizeof(T) == 2) {
            // x - tanh(x) cancels below 1/4: use the series x^3 (1/3 - 2 x^2/15 + 17 x^4/315) there
            const float x2 = x * x;
            const float small = x * x2 * fmaf(x2, fmaf(x2, 0.05396825397f, -0.13333333333f), 0.33333333333f);
            return fabsf(x) < 0.25f ? small : x - fast_tanh(x);
        }
        return x - tanhf(x);
    }
};

// Forward op of a continuous activation: y = fn(x), code = bucket(x), eight values at a time.
template <class Fn, typename T, int B> struct QuantizeOp {
    static constexpr int kBits = B;
    static constexpr bool kHeavy = true;  // transcendental math: see TileConfig in launch.cuh
    using Scratch = typename Bucketizer<T, B>::Scratch;
    Fn fn;
    Bucketizer<T, B> bucket;
    __device__ __forceinline__ void prepare(Scratch &s) { bucket.prepare(s); }
    __device__ __forceinline__ void apply(float (&v)[8], uint32_t (&code)[8]) const {
        bucket.lookup(v, code);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fn.template eval<T>(v[j]);
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
        return neg ? x + lambd : (pos ? x - lambd : 0.0f);
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
    __device__ __forceinline__ void apply(float (&v)[8], uint32_t (&code)[8]) const {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fn(v[j], code[j]);
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
