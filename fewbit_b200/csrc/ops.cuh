// fewbit_b200 -- per-element operators plugged into the warp-tile kernels of tile.cuh.
#pragma once

#include <math_constants.h>

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
//  B <= 3 : the (<= 7) bounds live in registers; a compare/select tree, B compares.
//  B >= 4 : bounds in shared memory; branch-free binary search, B-1 dependent LDS
//           (the first probe is a register).  For B >= 6 the table is skewed by one word
//           per 32 entries so that the 2^k probes of level k fall into distinct banks.
// Tables shorter than 2^B - 1 are padded with +inf (never counted).
// =====================================================================================

template <typename T, int B, bool kInRegisters = (B <= 3)> struct Bucketizer;

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

    __device__ __forceinline__ uint32_t operator()(float x) const {
        if constexpr (B == 1) {
            return b[0] < x ? 1u : 0u;
        } else if constexpr (B == 2) {
            const bool p1 = b[1] < x;
            const bool p2 = (p1 ? b[2] : b[0]) < x;
            return (p1 ? 2u : 0u) | (p2 ? 1u : 0u);
        } else {
            const bool p1 = b[3] < x;
            const bool p2 = (p1 ? b[5] : b[1]) < x;
            const float lo = p2 ? b[2] : b[0];
            const float hi = p2 ? b[6] : b[4];
            const bool p3 = (p1 ? hi : lo) < x;
            return (p1 ? 4u : 0u) | (p2 ? 2u : 0u) | (p3 ? 1u : 0u);
        }
    }
};

template <typename T, int B> struct Bucketizer<T, B, false> {
    static constexpr int kCount = (1 << B) - 1;
    static constexpr bool kSkew = B >= 6;
    static constexpr int skew(int i) { return kSkew ? i + (i >> 5) : i; }
    struct Scratch {
        float table[skew(kCount) + 1];
    };
    const T *bounds;
    int nbounds;
    const float *table;
    float top;

    __device__ __forceinline__ void prepare(Scratch &s) {
        for (int i = threadIdx.x; i < kCount; i += blockDim.x)
            s.table[skew(i)] = i < nbounds ? to_float<T>(bounds[i]) : CUDART_INF_F;
        __syncthreads();
        table = s.table;
        top = s.table[skew(kCount / 2)];
    }

    __device__ __forceinline__ uint32_t operator()(float x) const {
        constexpr int kTopStep = 1 << (B - 1);
        // `pos` is skew(idx) where idx counts the bounds known to be < x.
        int pos = top < x ? skew(kTopStep) : 0;
#pragma unroll
        for (int step = kTopStep >> 1; step >= 1; step >>= 1) {
            // probe entry idx + step - 1; its skewed address is pos + skew(step - 1), and
            // accepting it advances pos by skew(step) ("bank-skewed search").
            if (table[pos + skew(step - 1)] < x) pos += skew(step);
        }
        if constexpr (kSkew) pos -= (pos * 1986) >> 16;  // pos - pos / 33, valid for pos <= 262
        return (uint32_t)pos;
    }
};

// =====================================================================================
// Continuous activations.  Formulas are the ones ATen's CUDA kernels use in fp32 opmath
// (that is what the reference's test compares against: functional/activations_test.py:88-89),
// evaluated with the accurate libdevice functions; bf16 inputs are widened to fp32 and the
// result is rounded once.  Reference lambdas: fewbit/cuda/codec.cu:517-653.
// =====================================================================================

struct EluFamily {  // celu / elu / selu:  x > 0 ? x*pos : expm1(x*in_scale)*neg
    float pos, neg, in_scale;
    __device__ __forceinline__ float operator()(float x) const {
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
struct GeluFn {  // codec.cu:539-544 (x * normcdf(x)); ATen: x * 0.5 * (1 + erf(x / sqrt(2)))
    __host__ GeluFn(double, double) {}
    __device__ __forceinline__ float operator()(float x) const {
        return (x * 0.5f) * (1.0f + erff(x * 0.70710678118654752440f));
    }
};
struct HardswishFn {  // codec.cu:546-564
    __host__ HardswishFn(double, double) {}
    __device__ __forceinline__ float operator()(float x) const {
        return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
    }
};
struct LogSigmoidFn {  // codec.cu:566-576
    __host__ LogSigmoidFn(double, double) {}
    __device__ __forceinline__ float operator()(float x) const {
        return fminf(0.0f, x) - log1pf(expf(-fabsf(x)));
    }
};
struct MishFn {  // codec.cu:578-586
    __host__ MishFn(double, double) {}
    __device__ __forceinline__ float operator()(float x) const {
        return x * tanhf(log1pf(expf(x)));
    }
};
struct SigmoidFn {  // codec.cu:602-607
    __host__ SigmoidFn(double, double) {}
    __device__ __forceinline__ float operator()(float x) const {
        return 1.0f / (1.0f + expf(-x));
    }
};
struct SiluFn {  // codec.cu:609-614
    __host__ SiluFn(double, double) {}
    __device__ __forceinline__ float operator()(float x) const { return x / (1.0f + expf(-x)); }
};
struct SoftplusFn {  // codec.cu:616-632
    float beta, threshold;
    __host__ SoftplusFn(double b, double t) : beta((float)b), threshold((float)t) {}
    __device__ __forceinline__ float operator()(float x) const {
        const float bx = x * beta;
        return bx > threshold ? x : log1pf(expf(bx)) / beta;
    }
};
struct SoftsignFn {  // codec.cu:634-639
    __host__ SoftsignFn(double, double) {}
    __device__ __forceinline__ float operator()(float x) const { return x / (1.0f + fabsf(x)); }
};
struct TanhFn {  // codec.cu:641-646
    __host__ TanhFn(double, double) {}
    __device__ __forceinline__ float operator()(float x) const { return tanhf(x); }
};
struct TanhshrinkFn {  // codec.cu:648-653
    __host__ TanhshrinkFn(double, double) {}
    __device__ __forceinline__ float operator()(float x) const { return x - tanhf(x); }
};

// Forward op of a continuous activation: y = fn(x), code = bucket(x).
template <class Fn, typename T, int B> struct QuantizeOp {
    static constexpr int kBits = B;
    using Scratch = typename Bucketizer<T, B>::Scratch;
    Fn fn;
    Bucketizer<T, B> bucket;
    __device__ __forceinline__ void prepare(Scratch &s) { bucket.prepare(s); }
    __device__ __forceinline__ float apply(float x, uint32_t &code) const {
        code = bucket(x);
        return fn(x);
    }
};

// Backward op of every continuous activation: factor = levels[code]
// (StepwiseBackwardKernel, fewbit/cuda/codec.cu:655-670).
template <typename T, int B> struct LevelsOp {
    static constexpr int kBits = B;
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
    using Scratch = NoScratch;
    Fn fn;
    __device__ __forceinline__ void prepare(Scratch &) {}
    __device__ __forceinline__ float apply(float x, uint32_t &code) const { return fn(x, code); }
};

// Backward of the 1-bit family: factor = mask ? on : off  (codec.cu:271-296: idx * g;
// hardsigmoid :333-345: 1/6 | 0; leaky_relu :391-402: slope | 1).
struct MaskFactorOp {
    static constexpr int kBits = 1;
    using Scratch = NoScratch;
    float on, off;
    __device__ __forceinline__ void prepare(Scratch &) {}
    __device__ __forceinline__ float factor(uint32_t code) const { return code ? on : off; }
};

}  // namespace fewbit
