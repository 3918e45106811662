/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * fewbit_oracle: a scalar, single-threaded CPU restatement of the reference
 * algorithm for FewBit's quantized-gradient activation path.  It exists only
 * to check the CUDA product (fewbit_b200/) in tests/, in
 * __graft_entry__.smoke() and as bench.py's cpu_baseline leg.  Nothing under
 * fewbit_b200/ may import, link or execute it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 * against (a) the golden vectors held by the reference's own tests
 * (fewbit/cpu/codec_test.cc:12, fewbit/cuda/codec_test.cu:16-24,62-64,93-98),
 * (b) the unmodified reference CPU sources compiled into oracle/_ref/
 * (fewbit/cpu/codec.h, fewbit/cpu/gelu.cc) on random inputs, and (c) the
 * committed fixtures in tests/golden/ that were produced by running that
 * reference build (tests/golden/make_golden.py).
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the reference tree).
 */
#ifndef FEWBIT_ORACLE_H_
#define FEWBIT_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Continuous activation ids: order of CONTINOUS in fewbit/functional/activations.py:18-19. */
enum {
    ORC_CELU = 0, ORC_ELU, ORC_GELU, ORC_HARDSWISH, ORC_LOGSIGMOID, ORC_MISH, ORC_SELU,
    ORC_SIGMOID, ORC_SILU, ORC_SOFTPLUS, ORC_SOFTSIGN, ORC_TANH, ORC_TANHSHRINK,
    ORC_NUM_CONTINUOUS
};

/* Piecewise (1-bit) activation ids: order of STEPWISE in fewbit/functional/activations.py:14-15
 * (without 'stepwise', which has no kernel anywhere in the reference). */
enum {
    ORC_HARDSHRINK = 0, ORC_HARDSIGMOID, ORC_HARDTANH, ORC_LEAKY_RELU, ORC_RELU, ORC_RELU6,
    ORC_SOFTSHRINK, ORC_THRESHOLD,
    ORC_NUM_PIECEWISE
};

/* NaN policy of the bucket search. */
enum { ORC_NAN_TO_ZERO = 0 /* reference CUDA BinarySearch */, ORC_NAN_TO_LAST = 1 /* torch.searchsorted */ };

/* Number of bytes of the packed stream: ceil(n*bits/8) (fewbit/cpu/gelu.cc:18-20). */
size_t orc_state_bytes(int64_t n, int bits);

/* bits = ceil(log2(nlevels)), 1 for nlevels <= 2 (fewbit/cpu/gelu.cc:36, decision C-4). */
int orc_bits_for_levels(int nlevels);

/* LSB-first bit-stream packer / unpacker: fewbit/cpu/codec.h:33-57 and :59-83. */
void orc_deflate(const int32_t *codes, int64_t n, int bits, uint8_t *out);
void orc_inflate(int32_t *codes, int64_t n, int bits, const uint8_t *in);

/* code = #{i : bounds[i] < x}  == std::lower_bound == searchsorted(right=False):
 * fewbit/cuda/codec.cu:118-131, fewbit/cpu/gelu.cc:16. */
void orc_bucketize_f32(const float *x, int64_t n, const float *bounds, int nbounds,
                       int nan_policy, int32_t *codes);
/* bf16 carried as raw uint16 bit patterns. */
void orc_bucketize_bf16(const uint16_t *x, int64_t n, const uint16_t *bounds, int nbounds,
                        int nan_policy, int32_t *codes);

/* Fused forward of a continuous function (StepwiseKernel, fewbit/cuda/codec.cu:489-504;
 * Quantize, fewbit/cpu/gelu.cc:7-31): y = f(x) evaluated in double precision and rounded
 * once to the storage type (a correctly-rounded yardstick for the fp32 device math),
 * state = deflate(bucketize(x)).  p0/p1 = alpha | beta,threshold. */
void orc_stepwise_forward_f32(int func, const float *x, float *y, uint8_t *state, int64_t n,
                              int bits, const float *bounds, int nbounds, double p0, double p1,
                              int nan_policy);
void orc_stepwise_forward_bf16(int func, const uint16_t *x, uint16_t *y, uint8_t *state, int64_t n,
                               int bits, const uint16_t *bounds, int nbounds, double p0, double p1,
                               int nan_policy);

/* gin = levels[inflate(state)] * gout: StepwiseBackwardKernel fewbit/cuda/codec.cu:655-670,
 * QuantizeBackward fewbit/cpu/gelu.cc:33-45.  bf16: fp32 product, one RNE rounding. */
void orc_stepwise_backward_f32(const uint8_t *state, const float *gout, float *gin, int64_t n,
                               int bits, const float *levels);
void orc_stepwise_backward_bf16(const uint8_t *state, const uint16_t *gout, uint16_t *gin,
                                int64_t n, int bits, const uint16_t *levels);

/* 1-bit piecewise family, forward (value + mask) and backward:
 * fewbit/cuda/codec.cu:298-487 and the macro-generated backward kernels :271-296.
 * Decisions: relu6 saturates at 6.0 (C-6), pad bits are zero (C-7). */
void orc_piecewise_forward_f32(int func, const float *x, float *y, uint8_t *state, int64_t n,
                               double p0, double p1);
void orc_piecewise_forward_bf16(int func, const uint16_t *x, uint16_t *y, uint8_t *state,
                                int64_t n, double p0, double p1);
void orc_piecewise_backward_f32(int func, const uint8_t *state, const float *gout, float *gin,
                                int64_t n, double p0);
void orc_piecewise_backward_bf16(int func, const uint8_t *state, const uint16_t *gout,
                                 uint16_t *gin, int64_t n, double p0);

/* bf16 helpers (round-to-nearest-even, NaN preserved). */
uint16_t orc_f32_to_bf16(float v);
float orc_bf16_to_f32(uint16_t v);

#ifdef __cplusplus
}
#endif

#endif /* FEWBIT_ORACLE_H_ */
