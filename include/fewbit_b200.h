/*
 * fewbit_b200 -- C ABI of the B200-native FewBit hot path.
 *
 * This header is the drop-in boundary: plain pointers and sizes, no torch types.
 * Every entry point below replaces one launcher (or launcher family) of the reference's
 * CUDA layer; the torch operator library (fewbit_b200/libfewbit.so, namespace
 * `torch.ops.fewbit`) is a thin layer of C++ on top of exactly these calls.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers on the current CUDA device, borrowed for the
 *     duration of the enqueue (the caller keeps them alive until the stream has run).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - `x` and `y` may alias (the reference operators are in-place); `gout`/`gin` may alias.
 *   - Element counts are 64-bit (the reference is limited to uint32_t).
 *   - Packed state layout (identical to the reference CPU codec, fewbit/cpu/codec.h:33-57):
 *     element i occupies stream bits [i*bits, (i+1)*bits), LSB first inside little-endian
 *     bytes; length fewbit_state_bytes(n, bits) = ceil(n*bits/8); pad bits are zero.
 *   - Return value: 0 on success, >0 a cudaError_t from the launch, <0 a FEWBIT_E* code.
 *     Nothing here ever falls back to the CPU.
 */
#ifndef FEWBIT_B200_H_
#define FEWBIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FEWBIT_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define FEWBIT_API __attribute__((visibility("default")))
#else
#define FEWBIT_API
#endif

/* Storage type of x / y / gout / gin / bounds / levels (tables are passed in the
 * activation dtype, as the reference does: functional/activations.py:212). */
typedef enum { FEWBIT_F32 = 0, FEWBIT_BF16 = 1 } fewbit_dtype_t;

/* Continuous activations: order of CONTINOUS, fewbit/functional/activations.py:18-19;
 * one reference launcher each, fewbit/cuda/codec.h:80-92 (codec.cu:517-653). */
typedef enum {
    FEWBIT_CELU = 0, FEWBIT_ELU, FEWBIT_GELU, FEWBIT_HARDSWISH, FEWBIT_LOGSIGMOID, FEWBIT_MISH,
    FEWBIT_SELU, FEWBIT_SIGMOID, FEWBIT_SILU, FEWBIT_SOFTPLUS, FEWBIT_SOFTSIGN, FEWBIT_TANH,
    FEWBIT_TANHSHRINK, FEWBIT_NUM_CONTINUOUS
} fewbit_continuous_t;

/* Piecewise (1-bit mask) activations: fewbit/cuda/codec.h:59-68 (codec.cu:298-487). */
typedef enum {
    FEWBIT_HARDSHRINK = 0, FEWBIT_HARDSIGMOID, FEWBIT_HARDTANH, FEWBIT_LEAKY_RELU, FEWBIT_RELU,
    FEWBIT_RELU6, FEWBIT_SOFTSHRINK, FEWBIT_THRESHOLD, FEWBIT_NUM_PIECEWISE
} fewbit_piecewise_t;

enum {
    FEWBIT_OK = 0,
    FEWBIT_EINVAL = -1, /* null pointer, negative size, bits outside 1..8, bad table size */
    FEWBIT_EDTYPE = -2, /* unknown dtype */
    FEWBIT_EFUNC = -3,  /* unknown function id */
    FEWBIT_EALIGN = -4  /* element pointer not aligned to its element size */
};

FEWBIT_API int fewbit_abi_version(void);
FEWBIT_API const char *fewbit_error_string(int status);

/* ceil(n*bits/8).  Replaces the buffer sizing in ContinousCudaFunction::forward
 * (fewbit/cuda/activation.cc:349-351; there nobits = bits+1) and in
 * Quantize (fewbit/cpu/gelu.cc:18-19). */
FEWBIT_API size_t fewbit_state_bytes(int64_t n, int bits);

/* bits = ceil(log2(nlevels)), 1 for nlevels <= 2.  Replaces GetBitWidth/Log2
 * (fewbit/cuda/activation.cc:7-21) with the CPU op's rule (fewbit/cpu/gelu.cc:36). */
FEWBIT_API int fewbit_bits_for_levels(int nlevels);

/*
 * Fused forward of a continuous activation:  y = f(x);  code = #{i : bounds[i] < x};
 * state = pack(code, bits).  Replaces Celu..Tanhshrink (fewbit/cuda/codec.h:80-92) and
 * StepwiseKernel + BinarySearch + DeflateWarpKernel (fewbit/cuda/codec.cu:489-504,
 * 118-131, 142-165).
 *   bounds : nbounds <= 2^bits - 1 sorted interior borders, device memory, dtype `dtype`
 *   p0, p1 : alpha (celu, elu) | beta, threshold (softplus); ignored otherwise
 */
FEWBIT_API int fewbit_stepwise_forward(int func, int dtype, const void *x, void *y, uint8_t *state,
                            int64_t n, int bits, const void *bounds, int nbounds, double p0,
                            double p1, void *stream);

/* gin = levels[unpack(state)] * gout.  Replaces StepwiseBackward
 * (fewbit/cuda/codec.h:94-96, codec.cu:655-670 + InflateWarpKernel :184-203).
 *   levels : nlevels <= 2^bits values, device memory, dtype `dtype` */
FEWBIT_API int fewbit_stepwise_backward(int dtype, const uint8_t *state, const void *gout, void *gin,
                             int64_t n, int bits, const void *levels, int nlevels, void *stream);

/* Custom-table activation (reference schema `stepwise`, fewbit/fewbit.cc:37, and module
 * `Stepwise`, fewbit/modules/activations.py:97-134; the reference declares the operator and ships
 * no kernel for it).  The table is the function: `levels` (nlevels = nbounds + 1 values in the
 * activation dtype) are the slopes of a continuous piecewise-linear F with kinks at `bounds`,
 * F(anchor) = 0; y = F(x) (y may be x), state = pack(bucketize(x, bounds), bits) exactly as
 * fewbit_stepwise_forward.  Its backward is fewbit_stepwise_backward with the same `levels`. */
FEWBIT_API int fewbit_stepwise_custom_forward(int dtype, const void *x, void *y, uint8_t *state, int64_t n,
                                   int bits, const void *bounds, int nbounds, const void *levels,
                                   int nlevels, double anchor, void *stream);

/* 1-bit family forward: y = f(x), state = pack(mask, 1).  Replaces Hardshrink..Threshold and
 * LeakyRelu (fewbit/cuda/codec.h:59-68).
 *   p0, p1 : lambd | min_val,max_val | negative_slope | threshold,value */
FEWBIT_API int fewbit_piecewise_forward(int func, int dtype, const void *x, void *y, uint8_t *state,
                             int64_t n, double p0, double p1, void *stream);

/* 1-bit family backward: gin = factor(mask) * gout.  Replaces HardshrinkBackward ..
 * ThresholdBackward, LeakyReluBackward (fewbit/cuda/codec.h:59-68).  p0 = negative_slope. */
FEWBIT_API int fewbit_piecewise_backward(int func, int dtype, const uint8_t *state, const void *gout,
                              void *gin, int64_t n, double p0, void *stream);

/* Stand-alone codec on int32 codes (device memory).  Replaces DeflateBlock / InflateBlock
 * (fewbit/cuda/codec.h:18-26, codec.cu:166-220) with the exact-`bits` stream layout. */
FEWBIT_API int fewbit_deflate(const int32_t *codes, uint8_t *state, int64_t n, int bits, void *stream);
FEWBIT_API int fewbit_inflate(const uint8_t *state, int32_t *codes, int64_t n, int bits, void *stream);

/*
 * Host-buffer variants (the "e2e" path of bench.py): x / y / gout / gin are HOST pointers
 * (pinned for full speed); the call stages them through device memory in chunks on three
 * streams so that H2D, kernel and D2H overlap, and returns after the last byte is back.
 * `state` stays a DEVICE pointer (it is the tensor saved for backward).  Tables are device
 * pointers as above.  `chunk_elems` <= 0 picks a default.
 *
 * Ordering: the staging streams are private to the library and not ordered against any stream of
 * the caller.  Whatever produced `state` / the tables must have completed on the device before the
 * call (e.g. cudaStreamSynchronize on the producing stream), and the call returns only after all
 * of its own device work has completed -- on success AND on failure -- so the host buffers and
 * `state` may be reused immediately.  One staging pipeline exists per device; concurrent calls for
 * the same device are serialised, calls for different devices run side by side.
 */
FEWBIT_API int fewbit_stepwise_forward_host(int func, int dtype, const void *x_host, void *y_host,
                                 uint8_t *state, int64_t n, int bits, const void *bounds,
                                 int nbounds, double p0, double p1, int64_t chunk_elems);
FEWBIT_API int fewbit_stepwise_backward_host(int dtype, const uint8_t *state, const void *gout_host,
                                  void *gin_host, int64_t n, int bits, const void *levels,
                                  int nlevels, int64_t chunk_elems);
FEWBIT_API int fewbit_piecewise_forward_host(int func, int dtype, const void *x_host, void *y_host,
                                  uint8_t *state, int64_t n, double p0, double p1,
                                  int64_t chunk_elems);
FEWBIT_API int fewbit_piecewise_backward_host(int func, int dtype, const uint8_t *state,
                                   const void *gout_host, void *gin_host, int64_t n, double p0,
                                   int64_t chunk_elems);

/*
 * RandomizedLinear's projection on tcgen05 tensor cores:  out[P, D] = scale * S[P, N] X[N, D]
 * with the random sketch S generated inside the kernel (never materialised): entry (p, n) is a
 * pure function of (seed, offset, p, n), so forward and backward see the same S.  Replaces
 * `proj = randn(P, N); (proj @ input_view) / P` and `proj @ grad_output_view` of
 * LinearGRPFunc (fewbit/functional/linear.py:133-137, 196-199).
 *   x         : [tokens, features] bf16, row-major, 16-byte aligned, features % 8 == 0
 *   out       : [rows, features] fp32
 *   workspace : fewbit_sketch_workspace_bytes(...) bytes of device memory (split-K partials;
 *               may be NULL when that is 0)
 *   kind      : 0 = N(0,1) entries ('gaussian'), 1 = +-1/2 entries ('rademacher')
 * fewbit_sketch_project is the same product with the two passes that follow it in
 * LinearGRPFunc.backward folded in (fewbit/functional/linear.py:199-217):
 *   out_dtype   : FEWBIT_F32 or FEWBIT_BF16 -- the result is rounded once, in the kernel, instead of
 *                 by a separate `.to(dtype)` pass over [P, D];
 *   column_sums : non-zero appends one row of ones to S: out has rows + 1 rows and the last one is
 *                 scale * sum_n X[n, :], i.e. `grad_output.sum(0)` (the bias gradient) for the price
 *                 of one more sketch row.  Size the workspace for rows + 1.
 * fewbit_sketch_matrix writes S itself ([rows, cols] bf16) -- for tests and diagnostics only.
 */
FEWBIT_API size_t fewbit_sketch_workspace_bytes(int64_t tokens, int features, int rows);
FEWBIT_API int fewbit_sketch_forward(const void *x, float *out, void *workspace, int64_t tokens,
                                     int features, int rows, int kind, float scale, uint64_t seed,
                                     uint64_t offset, void *stream);
FEWBIT_API int fewbit_sketch_project(const void *x, void *out, int out_dtype, void *workspace,
                                     int64_t tokens, int features, int rows, int column_sums, int kind,
                                     float scale, uint64_t seed, uint64_t offset, void *stream);
FEWBIT_API int fewbit_sketch_matrix(void *s_bf16, int rows, int64_t cols, int kind, uint64_t seed,
                                    uint64_t offset, void *stream);
/* The launch plan fewbit_sketch_forward would use for a shape on a GPU with `sms` multiprocessors (no
 * device needed; host tests check its invariants over many shapes).  plan[8] = { BN (sketch rows per CTA),
 * split_k, CTAs sharing one generated S slot, 1 if those are a cta_group::2 pair, 64-token stages per
 * split, S ring slots, bytes of one 64-token S tile, dynamic shared memory in bytes }. */
FEWBIT_API int fewbit_sketch_plan(int64_t tokens, int features, int rows, int kind, int sms, int plan[8]);

/* Number of kernels this library has launched in the calling process (for bench.py's
 * `gpu_launches`). */
FEWBIT_API int64_t fewbit_launch_count(void);

#ifdef __cplusplus
}
#endif

#endif /* FEWBIT_B200_H_ */
