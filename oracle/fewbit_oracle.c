/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See fewbit_oracle.h for the contract.
 *
 * Scalar CPU restatement of the reference's quantized-gradient activation path.
 * Deliberately naive: one element at a time, one bit cursor, libm in double.
 * Parity status: PINNED (tests/test_oracle.py, tests/golden/).
 */
#include "fewbit_oracle.h"

#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------ bf16 -- */

float orc_bf16_to_f32(uint16_t v) {
    uint32_t u = (uint32_t)v << 16;
    float f;
    memcpy(&f, &u, sizeof f);
    return f;
}

uint16_t orc_f32_to_bf16(float v) {
    uint32_t u;
    memcpy(&u, &v, sizeof u);
    if ((u & 0x7fffffffu) > 0x7f800000u) { /* NaN: keep sign, force a quiet payload */
        return (uint16_t)((u >> 16) | 0x0040u);
    }
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb; /* round to nearest, ties to even */
    return (uint16_t)(u >> 16);
}

/* ----------------------------------------------------------- stream size -- */

size_t orc_state_bytes(int64_t n, int bits) {
    if (n <= 0) return 0;
    return (size_t)(((uint64_t)n * (uint64_t)bits + 7u) / 8u);
}

int orc_bits_for_levels(int nlevels) {
    int bits = 1;
    while ((1 << bits) < nlevels) ++bits;
    return bits;
}

/* ------------------------------------------------------------------ codec -- */

/* Reference: fewbit/cpu/codec.h:33-57 (Deflate<uint8_t>).  Element i occupies
 * stream bits [i*bits, (i+1)*bits), least-significant bit first inside
 * little-endian bytes; unused bits of the last byte are zero. */
void orc_deflate(const int32_t *codes, int64_t n, int bits, uint8_t *out) {
    size_t nbytes = orc_state_bytes(n, bits);
    memset(out, 0, nbytes);
    uint32_t keep = (bits >= 32) ? 0xffffffffu : ((1u << bits) - 1u);
    for (int64_t i = 0; i < n; ++i) {
        uint64_t cursor = (uint64_t)i * (uint64_t)bits;
        uint32_t value = (uint32_t)codes[i] & keep;
        int remaining = bits;
        while (remaining > 0) {
            size_t byte = (size_t)(cursor >> 3);
            int offset = (int)(cursor & 7u);
            int room = 8 - offset;
            int take = remaining < room ? remaining : room;
            out[byte] |= (uint8_t)((value & ((1u << take) - 1u)) << offset);
            value >>= take;
            cursor += (uint64_t)take;
            remaining -= take;
        }
    }
}

/* Reference: fewbit/cpu/codec.h:59-83 (Inflate<uint8_t>). */
void orc_inflate(int32_t *codes, int64_t n, int bits, const uint8_t *in) {
    for (int64_t i = 0; i < n; ++i) {
        uint64_t cursor = (uint64_t)i * (uint64_t)bits;
        uint32_t value = 0;
        int got = 0;
        while (got < bits) {
            size_t byte = (size_t)(cursor >> 3);
            int offset = (int)(cursor & 7u);
            int room = 8 - offset;
            int take = (bits - got) < room ? (bits - got) : room;
            uint32_t piece = ((uint32_t)in[byte] >> offset) & ((1u << take) - 1u);
            value |= piece << got;
            got += take;
            cursor += (uint64_t)take;
        }
        codes[i] = (int32_t)value;
    }
}

/* -------------------------------------------------------------- bucketize -- */

/* Reference: BinarySearch fewbit/cuda/codec.cu:118-131 (std::lower_bound: first i with
 * !(bounds[i] < x)); torch::searchsorted(bounds, x, right=False) fewbit/cpu/gelu.cc:16.
 * On sorted bounds both equal the strict count below.  They differ only on NaN:
 * every `bounds[i] < NaN` is false -> 0 (CUDA), searchsorted sorts NaN last -> nbounds. */
static int32_t bucket_of(float x, const float *bounds, int nbounds, int nan_policy) {
    if (x != x) return nan_policy == ORC_NAN_TO_LAST ? nbounds : 0;
    int32_t count = 0;
    for (int i = 0; i < nbounds; ++i) count += (bounds[i] < x) ? 1 : 0;
    return count;
}

void orc_bucketize_f32(const float *x, int64_t n, const float *bounds, int nbounds,
                       int nan_policy, int32_t *codes) {
    for (int64_t i = 0; i < n; ++i) codes[i] = bucket_of(x[i], bounds, nbounds, nan_policy);
}

void orc_bucketize_bf16(const uint16_t *x, int64_t n, const uint16_t *bounds, int nbounds,
                        int nan_policy, int32_t *codes) {
    float fb[256];
    if (nbounds > 256) nbounds = 256;
    for (int i = 0; i < nbounds; ++i) fb[i] = orc_bf16_to_f32(bounds[i]);
    for (int64_t i = 0; i < n; ++i)
        codes[i] = bucket_of(orc_bf16_to_f32(x[i]), fb, nbounds, nan_policy);
}

/* ------------------------------------------------- continuous activations -- */

/* Forward values, evaluated in double on the (float-cast) parameters and rounded once.
 * Formulas: fewbit/cuda/codec.cu:517-653 (the same functions torch.nn.functional.* define,
 * which is what the reference's own test compares against, functional/activations_test.py:88-89). */
static double eval_continuous(int func, double x, double p0, double p1) {
    const double selu_alpha = 1.6732632423543772848170429916717;
    const double selu_scale = 1.0507009873554804934193349852946;
    switch (func) {
    case ORC_CELU: { /* codec.cu:517-526 */
        double a = (double)(float)p0;
        return x > 0 ? x : a * expm1(x / a);
    }
    case ORC_ELU: { /* codec.cu:528-537 */
        double a = (double)(float)p0;
        return x > 0 ? x : a * expm1(x);
    }
    case ORC_GELU: /* codec.cu:539-544: x * Phi(x) */
        return 0.5 * x * erfc(-x * 0.70710678118654752440);
    case ORC_HARDSWISH: { /* codec.cu:546-564 */
        double t = x + 3.0;
        t = t < 0 ? 0 : (t > 6 ? 6 : t);
        return x * t / 6.0;
    }
    case ORC_LOGSIGMOID: /* codec.cu:566-576 */
        return (x < 0 ? x : 0.0) - log1p(exp(-fabs(x)));
    case ORC_MISH: /* codec.cu:578-586 */
        return x * tanh(log1p(exp(x)));
    case ORC_SELU: /* codec.cu:588-600 */
        return x > 0 ? (double)(float)selu_scale * x
                     : (double)(float)selu_scale * (double)(float)selu_alpha * expm1(x);
    case ORC_SIGMOID: /* codec.cu:602-607 */
        return 1.0 / (1.0 + exp(-x));
    case ORC_SILU: /* codec.cu:609-614 */
        return x / (1.0 + exp(-x));
    case ORC_SOFTPLUS: { /* codec.cu:616-632 */
        double beta = (double)(float)p0, thr = (double)(float)p1;
        return beta * x > thr ? x : log1p(exp(beta * x)) / beta;
    }
    case ORC_SOFTSIGN: /* codec.cu:634-639 */
        return x / (1.0 + fabs(x));
    case ORC_TANH: /* codec.cu:641-646 */
        return tanh(x);
    case ORC_TANHSHRINK: /* codec.cu:648-653 */
        return x - tanh(x);
    default:
        return NAN;
    }
}

void orc_stepwise_forward_f32(int func, const float *x, float *y, uint8_t *state, int64_t n,
                              int bits, const float *bounds, int nbounds, double p0, double p1,
                              int nan_policy) {
    size_t nbytes = orc_state_bytes(n, bits);
    memset(state, 0, nbytes);
    for (int64_t i = 0; i < n; ++i) {
        float xi = x[i]; /* read before write: y may alias x (in-place op) */
        int32_t code = bucket_of(xi, bounds, nbounds, nan_policy);
        y[i] = (float)eval_continuous(func, (double)xi, p0, p1);
        uint64_t cursor = (uint64_t)i * (uint64_t)bits;
        uint32_t wide = (uint32_t)code << (cursor & 7u); /* bits <= 8: spans <= 2 bytes */
        state[cursor >> 3] |= (uint8_t)wide;
        if ((wide >> 8) != 0) state[(cursor >> 3) + 1] |= (uint8_t)(wide >> 8);
    }
}

void orc_stepwise_forward_bf16(int func, const uint16_t *x, uint16_t *y, uint8_t *state, int64_t n,
                               int bits, const uint16_t *bounds, int nbounds, double p0, double p1,
                               int nan_policy) {
    float fb[256];
    if (nbounds > 256) nbounds = 256;
    for (int i = 0; i < nbounds; ++i) fb[i] = orc_bf16_to_f32(bounds[i]);
    size_t nbytes = orc_state_bytes(n, bits);
    memset(state, 0, nbytes);
    for (int64_t i = 0; i < n; ++i) {
        float xi = orc_bf16_to_f32(x[i]);
        int32_t code = bucket_of(xi, fb, nbounds, nan_policy);
        /* double -> float -> bf16 would round twice; go through float only when it is
         * exact enough: the device computes in fp32 and rounds once, so mirror that. */
        y[i] = orc_f32_to_bf16((float)eval_continuous(func, (double)xi, p0, p1));
        uint64_t cursor = (uint64_t)i * (uint64_t)bits;
        uint32_t wide = (uint32_t)code << (cursor & 7u);
        state[cursor >> 3] |= (uint8_t)wide;
        if ((wide >> 8) != 0) state[(cursor >> 3) + 1] |= (uint8_t)(wide >> 8);
    }
}

/* Reference: StepwiseBackwardKernel fewbit/cuda/codec.cu:655-670; QuantizeBackward
 * fewbit/cpu/gelu.cc:33-45 (levels.index(codes) * grads). */
static uint32_t code_at(const uint8_t *state, int64_t i, int bits) {
    uint64_t cursor = (uint64_t)i * (uint64_t)bits;
    size_t byte = (size_t)(cursor >> 3);
    int offset = (int)(cursor & 7u);
    uint32_t window = state[byte];
    if (offset + bits > 8) window |= (uint32_t)state[byte + 1] << 8;
    return (window >> offset) & ((1u << bits) - 1u);
}

void orc_stepwise_backward_f32(const uint8_t *state, const float *gout, float *gin, int64_t n,
                               int bits, const float *levels) {
    for (int64_t i = 0; i < n; ++i) gin[i] = levels[code_at(state, i, bits)] * gout[i];
}

void orc_stepwise_backward_bf16(const uint8_t *state, const uint16_t *gout, uint16_t *gin,
                                int64_t n, int bits, const uint16_t *levels) {
    for (int64_t i = 0; i < n; ++i) {
        float lv = orc_bf16_to_f32(levels[code_at(state, i, bits)]);
        gin[i] = orc_f32_to_bf16(lv * orc_bf16_to_f32(gout[i])); /* exact product, one rounding */
    }
}

/* ---------------------------------------------------- piecewise (1 bit) -- */

/* One element of the eight forward kernels, branch order as in the reference so that NaN
 * inputs land in the same branch.  `round_param` rounds a parameter that is written to the
 * output to the storage type (identity for fp32). */
static float piecewise_value(int func, float x, float p0, float p1, int *mask) {
    switch (func) {
    case ORC_HARDSHRINK: /* codec.cu:298-311 */
        if (x < -p0 || x > p0) { *mask = 1; return x; }
        *mask = 0; return 0.0f;
    case ORC_HARDSIGMOID: /* codec.cu:316-331; value as F.hardsigmoid: min(max(x+3,0),6)/6 */
        if (x <= -3.0f) { *mask = 0; return 0.0f; }
        if (x >= 3.0f) { *mask = 0; return 1.0f; }
        *mask = 1; return (float)(((double)x + 3.0) / 6.0);
    case ORC_HARDTANH: /* codec.cu:354-370 */
        if (x <= p0) { *mask = 0; return p0; }
        if (x >= p1) { *mask = 0; return p1; }
        *mask = 1; return x;
    case ORC_LEAKY_RELU: /* codec.cu:375-389 : mask marks the NEGATIVE side */
        if (x >= 0.0f) { *mask = 0; return x; }
        *mask = 1; return p0 * x;
    case ORC_RELU: /* codec.cu:412-425 */
        if (x <= 0.0f) { *mask = 0; return 0.0f; }
        *mask = 1; return x;
    case ORC_RELU6: /* codec.cu:430-445; saturates at 6.0 (decision C-6, reference writes 1.0) */
        if (x <= 0.0f) { *mask = 0; return 0.0f; }
        if (x >= 6.0f) { *mask = 0; return 6.0f; }
        *mask = 1; return x;
    case ORC_SOFTSHRINK: /* codec.cu:450-465 */
        if (x < -p0) { *mask = 1; return x + p0; }
        if (x > p0) { *mask = 1; return x - p0; }
        *mask = 0; return 0.0f;
    case ORC_THRESHOLD: /* codec.cu:470-484 */
        if (x <= p0) { *mask = 0; return p1; }
        *mask = 1; return x;
    default:
        *mask = 0; return NAN;
    }
}

/* Backward multiplier: macro kernels codec.cu:271-296 (idx * g), hardsigmoid :333-345
 * (1/6 or 0), leaky_relu :391-402 (slope or 1). */
static float piecewise_factor(int func, int mask, float p0) {
    switch (func) {
    case ORC_HARDSIGMOID: return mask ? 1.0f / 6.0f : 0.0f;
    case ORC_LEAKY_RELU: return mask ? p0 : 1.0f;
    default: return mask ? 1.0f : 0.0f;
    }
}

void orc_piecewise_forward_f32(int func, const float *x, float *y, uint8_t *state, int64_t n,
                               double p0, double p1) {
    memset(state, 0, orc_state_bytes(n, 1));
    for (int64_t i = 0; i < n; ++i) {
        int mask;
        y[i] = piecewise_value(func, x[i], (float)p0, (float)p1, &mask);
        state[i >> 3] |= (uint8_t)(mask << (i & 7));
    }
}

void orc_piecewise_forward_bf16(int func, const uint16_t *x, uint16_t *y, uint8_t *state,
                                int64_t n, double p0, double p1) {
    memset(state, 0, orc_state_bytes(n, 1));
    for (int64_t i = 0; i < n; ++i) {
        int mask;
        float v = piecewise_value(func, orc_bf16_to_f32(x[i]), (float)p0, (float)p1, &mask);
        y[i] = orc_f32_to_bf16(v);
        state[i >> 3] |= (uint8_t)(mask << (i & 7));
    }
}

void orc_piecewise_backward_f32(int func, const uint8_t *state, const float *gout, float *gin,
                                int64_t n, double p0) {
    for (int64_t i = 0; i < n; ++i) {
        int mask = (state[i >> 3] >> (i & 7)) & 1;
        gin[i] = piecewise_factor(func, mask, (float)p0) * gout[i];
    }
}

void orc_piecewise_backward_bf16(int func, const uint8_t *state, const uint16_t *gout,
                                 uint16_t *gin, int64_t n, double p0) {
    for (int64_t i = 0; i < n; ++i) {
        int mask = (state[i >> 3] >> (i & 7)) & 1;
        gin[i] = orc_f32_to_bf16(piecewise_factor(func, mask, (float)p0) *
                                 orc_bf16_to_f32(gout[i]));
    }
}
