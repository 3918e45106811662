// fewbit_b200 -- the C ABI (include/fewbit_b200.h): validation, dispatch, host-staged variants.
#include <atomic>
#include <mutex>

#include "../../include/fewbit_b200.h"
#include "launch.cuh"

namespace fewbit {

// One launcher per continuous activation, each defined in its own translation unit
// (continuous_inst.cu compiled with -DFEWBIT_FN/-DFEWBIT_ENTRY).
#define FEWBIT_DECLARE(name) cudaError_t launch_forward_##name(const ForwardArgs &);
FEWBIT_DECLARE(celu) FEWBIT_DECLARE(elu) FEWBIT_DECLARE(gelu) FEWBIT_DECLARE(hardswish)
FEWBIT_DECLARE(logsigmoid) FEWBIT_DECLARE(mish) FEWBIT_DECLARE(selu) FEWBIT_DECLARE(sigmoid)
FEWBIT_DECLARE(silu) FEWBIT_DECLARE(softplus) FEWBIT_DECLARE(softsign) FEWBIT_DECLARE(tanh)
FEWBIT_DECLARE(tanhshrink)
#undef FEWBIT_DECLARE

cudaError_t launch_levels_backward(const BackwardArgs &);
cudaError_t launch_custom_forward(const ForwardArgs &, const void *levels, int nlevels);
cudaError_t launch_piecewise_forward(int func, const ForwardArgs &);
cudaError_t launch_piecewise_backward(int func, const BackwardArgs &);
cudaError_t launch_deflate(const int32_t *, uint8_t *, int64_t, int, cudaStream_t);
cudaError_t launch_inflate(const uint8_t *, int32_t *, int64_t, int, cudaStream_t);

static std::atomic<int64_t> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

using ForwardLauncher = cudaError_t (*)(const ForwardArgs &);
static const ForwardLauncher kContinuous[FEWBIT_NUM_CONTINUOUS] = {
    launch_forward_celu,    launch_forward_elu,      launch_forward_gelu,
    launch_forward_hardswish, launch_forward_logsigmoid, launch_forward_mish,
    launch_forward_selu,    launch_forward_sigmoid,  launch_forward_silu,
    launch_forward_softplus, launch_forward_softsign, launch_forward_tanh,
    launch_forward_tanhshrink,
};

static int elem_size(int dtype) { return dtype == FEWBIT_F32 ? 4 : 2; }

static int check_common(int dtype, const void *a, const void *b, const void *state, int64_t n) {
    if (dtype != FEWBIT_F32 && dtype != FEWBIT_BF16) return FEWBIT_EDTYPE;
    if (n < 0) return FEWBIT_EINVAL;
    if (n == 0) return FEWBIT_OK;
    if (!a || !b || !state) return FEWBIT_EINVAL;
    const uintptr_t m = (uintptr_t)elem_size(dtype) - 1;
    if (((uintptr_t)a & m) || ((uintptr_t)b & m)) return FEWBIT_EALIGN;
    return FEWBIT_OK;
}

// -------------------------------------------------------------------------------------
// Host-staged pipeline: chunks of the host arrays travel H2D -> kernel -> D2H on a ring of
// streams, so the two PCIe directions and the kernel overlap.
// -------------------------------------------------------------------------------------
struct HostPipeline {
    static constexpr int kSlots = 3;
    std::mutex mu;
    cudaStream_t streams[kSlots] = {};
    void *in[kSlots] = {}, *out[kSlots] = {};
    size_t capacity = 0;

    // Called with `mu` held and this pipeline's device current.
    cudaError_t ensure(size_t bytes) {
        if (bytes <= capacity) return cudaSuccess;
        release();
        cudaError_t e;
        for (int i = 0; i < kSlots; ++i) {
            if ((e = cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&in[i], bytes)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&out[i], bytes)) != cudaSuccess) return e;
        }
        capacity = bytes;
        return cudaSuccess;
    }
    void release() {
        for (int i = 0; i < kSlots; ++i) {
            if (in[i]) cudaFree(in[i]);
            if (out[i]) cudaFree(out[i]);
            if (streams[i]) cudaStreamDestroy(streams[i]);
            in[i] = out[i] = nullptr;
            streams[i] = nullptr;
        }
        capacity = 0;
    }
    // Nothing may still be reading or writing the caller's host buffers when a *_host call returns,
    // whether it succeeded or not.
    cudaError_t drain() {
        cudaError_t first = cudaSuccess;
        for (int i = 0; i < kSlots; ++i)
            if (streams[i]) {
                const cudaError_t e = cudaStreamSynchronize(streams[i]);
                if (first == cudaSuccess) first = e;
            }
        return first;
    }
};

// One pipeline per device (staging buffers and streams belong to a device); calls for different
// devices do not serialise against each other, calls for the same device do.
static HostPipeline *pipeline_of_current_device(cudaError_t *err) {
    static HostPipeline pipes[64];
    int dev = 0;
    *err = cudaGetDevice(&dev);
    if (*err != cudaSuccess) return nullptr;
    if (dev < 0 || dev >= 64) {
        *err = cudaErrorInvalidDevice;
        return nullptr;
    }
    return &pipes[dev];
}

// `enqueue(dev_in, dev_out, first_elem, count, stream)` launches the kernel for one chunk.
// The ring streams are private and non-blocking: the device state buffer and the tables are read
// and written on them with no ordering against the caller's own streams, so the caller must not
// have work in flight on those buffers (see include/fewbit_b200.h).
template <class Enqueue>
static int run_host_pipeline(int dtype, const void *src_host, void *dst_host, int64_t n,
                             int64_t chunk_elems, Enqueue enqueue) {
    if (n == 0) return FEWBIT_OK;
    const int es = elem_size(dtype);
    if (chunk_elems <= 0) chunk_elems = (int64_t)1 << 24;        // 16 Mi elements
    chunk_elems = std::max<int64_t>(2048, (chunk_elems / 2048) * 2048);  // whole warp tiles, 16 B state
    chunk_elems = std::min<int64_t>(chunk_elems, ((n + 2047) / 2048) * 2048);
    cudaError_t e = cudaSuccess;
    HostPipeline *pipe = pipeline_of_current_device(&e);
    if (!pipe) return (int)e;
    std::lock_guard<std::mutex> lock(pipe->mu);
    if ((e = pipe->ensure((size_t)chunk_elems * es)) != cudaSuccess) {
        pipe->release();
        return (int)e;
    }
    int status = FEWBIT_OK, slot = 0;
    for (int64_t first = 0; first < n && status == FEWBIT_OK;
         first += chunk_elems, slot = (slot + 1) % HostPipeline::kSlots) {
        const int64_t count = std::min(chunk_elems, n - first);
        cudaStream_t s = pipe->streams[slot];
        // the slot's previous D2H must be done before its buffers are reused: same stream -> ordered
        e = cudaMemcpyAsync(pipe->in[slot], (const char *)src_host + first * es, (size_t)count * es,
                            cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { status = (int)e; break; }
        status = enqueue(pipe->in[slot], pipe->out[slot], first, count, s);
        if (status != FEWBIT_OK) break;
        e = cudaMemcpyAsync((char *)dst_host + first * es, pipe->out[slot], (size_t)count * es,
                            cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) status = (int)e;
    }
    e = pipe->drain();        // also on the error path: earlier chunks are still in flight
    if (status == FEWBIT_OK && e != cudaSuccess) status = (int)e;
    return status;
}

}  // namespace fewbit

using namespace fewbit;

extern "C" {

int fewbit_abi_version(void) { return FEWBIT_B200_ABI_VERSION; }

const char *fewbit_error_string(int status) {
    switch (status) {
        case FEWBIT_OK: return "success";
        case FEWBIT_EINVAL: return "fewbit: invalid argument (null pointer, negative size, bits outside 1..8 or table too large)";
        case FEWBIT_EDTYPE: return "fewbit: unsupported dtype (expected FEWBIT_F32 or FEWBIT_BF16)";
        case FEWBIT_EFUNC: return "fewbit: unknown activation id";
        case FEWBIT_EALIGN: return "fewbit: element pointer is not aligned to its element size";
        default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "fewbit: unknown status";
    }
}

size_t fewbit_state_bytes(int64_t n, int bits) {
    if (n <= 0 || bits <= 0) return 0;
    return (size_t)(((uint64_t)n * (uint64_t)bits + 7u) / 8u);
}

int fewbit_bits_for_levels(int nlevels) {
    int bits = 1;
    while ((1 << bits) < nlevels) ++bits;
    return bits;
}

int fewbit_stepwise_forward(int func, int dtype, const void *x, void *y, uint8_t *state, int64_t n,
                            int bits, const void *bounds, int nbounds, double p0, double p1,
                            void *stream) {
    if (func < 0 || func >= FEWBIT_NUM_CONTINUOUS) return FEWBIT_EFUNC;
    if (bits < 1 || bits > 8 || nbounds < 0 || nbounds > (1 << bits) - 1) return FEWBIT_EINVAL;
    if (int st = check_common(dtype, x, y, state, n)) return st;
    if (n == 0) return FEWBIT_OK;
    if (nbounds > 0 && !bounds) return FEWBIT_EINVAL;
    ForwardArgs a{dtype, x, y, state, n, bits, bounds, nbounds, p0, p1, (cudaStream_t)stream};
    return (int)kContinuous[func](a);
}

int fewbit_stepwise_custom_forward(int dtype, const void *x, void *y, uint8_t *state, int64_t n, int bits,
                                   const void *bounds, int nbounds, const void *levels, int nlevels,
                                   double anchor, void *stream) {
    if (bits < 1 || bits > 8 || nbounds < 0 || nbounds > (1 << bits) - 1 || nlevels != nbounds + 1)
        return FEWBIT_EINVAL;
    if (int st = check_common(dtype, x, y, state, n)) return st;
    if (n == 0) return FEWBIT_OK;
    if ((nbounds > 0 && !bounds) || !levels) return FEWBIT_EINVAL;
    ForwardArgs a{dtype, x, y, state, n, bits, bounds, nbounds, anchor, 0.0, (cudaStream_t)stream};
    return (int)launch_custom_forward(a, levels, nlevels);
}

int fewbit_stepwise_backward(int dtype, const uint8_t *state, const void *gout, void *gin, int64_t n,
                             int bits, const void *levels, int nlevels, void *stream) {
    if (bits < 1 || bits > 8 || nlevels < 1 || nlevels > (1 << bits)) return FEWBIT_EINVAL;
    if (int st = check_common(dtype, gout, gin, state, n)) return st;
    if (n == 0) return FEWBIT_OK;
    if (!levels) return FEWBIT_EINVAL;
    BackwardArgs a{dtype, state, gout, gin, n, bits, levels, nlevels, 0.0, (cudaStream_t)stream};
    return (int)launch_levels_backward(a);
}

int fewbit_piecewise_forward(int func, int dtype, const void *x, void *y, uint8_t *state, int64_t n,
                             double p0, double p1, void *stream) {
    if (func < 0 || func >= FEWBIT_NUM_PIECEWISE) return FEWBIT_EFUNC;
    if (int st = check_common(dtype, x, y, state, n)) return st;
    if (n == 0) return FEWBIT_OK;
    ForwardArgs a{dtype, x, y, state, n, 1, nullptr, 0, p0, p1, (cudaStream_t)stream};
    return (int)launch_piecewise_forward(func, a);
}

int fewbit_piecewise_backward(int func, int dtype, const uint8_t *state, const void *gout, void *gin,
                              int64_t n, double p0, void *stream) {
    if (func < 0 || func >= FEWBIT_NUM_PIECEWISE) return FEWBIT_EFUNC;
    if (int st = check_common(dtype, gout, gin, state, n)) return st;
    if (n == 0) return FEWBIT_OK;
    BackwardArgs a{dtype, state, gout, gin, n, 1, nullptr, 0, p0, (cudaStream_t)stream};
    return (int)launch_piecewise_backward(func, a);
}

int fewbit_deflate(const int32_t *codes, uint8_t *state, int64_t n, int bits, void *stream) {
    if (bits < 1 || bits > 8 || n < 0) return FEWBIT_EINVAL;
    if (n == 0) return FEWBIT_OK;
    if (!codes || !state) return FEWBIT_EINVAL;
    return (int)launch_deflate(codes, state, n, bits, (cudaStream_t)stream);
}

int fewbit_inflate(const uint8_t *state, int32_t *codes, int64_t n, int bits, void *stream) {
    if (bits < 1 || bits > 8 || n < 0) return FEWBIT_EINVAL;
    if (n == 0) return FEWBIT_OK;
    if (!codes || !state) return FEWBIT_EINVAL;
    return (int)launch_inflate(state, codes, n, bits, (cudaStream_t)stream);
}

int fewbit_stepwise_forward_host(int func, int dtype, const void *x_host, void *y_host, uint8_t *state,
                                 int64_t n, int bits, const void *bounds, int nbounds, double p0,
                                 double p1, int64_t chunk_elems) {
    if (func < 0 || func >= FEWBIT_NUM_CONTINUOUS) return FEWBIT_EFUNC;
    if (bits < 1 || bits > 8 || nbounds < 0 || nbounds > (1 << bits) - 1) return FEWBIT_EINVAL;
    if (int st = check_common(dtype, x_host, y_host, state, n)) return st;
    return run_host_pipeline(dtype, x_host, y_host, n, chunk_elems,
                             [&](void *din, void *dout, int64_t first, int64_t count, cudaStream_t s) {
                                 return fewbit_stepwise_forward(func, dtype, din, dout,
                                                                state + first / 8 * bits, count, bits,
                                                                bounds, nbounds, p0, p1, s);
                             });
}

int fewbit_stepwise_backward_host(int dtype, const uint8_t *state, const void *gout_host, void *gin_host,
                                  int64_t n, int bits, const void *levels, int nlevels,
                                  int64_t chunk_elems) {
    if (bits < 1 || bits > 8 || nlevels < 1 || nlevels > (1 << bits)) return FEWBIT_EINVAL;
    if (int st = check_common(dtype, gout_host, gin_host, state, n)) return st;
    return run_host_pipeline(dtype, gout_host, gin_host, n, chunk_elems,
                             [&](void *din, void *dout, int64_t first, int64_t count, cudaStream_t s) {
                                 return fewbit_stepwise_backward(dtype, state + first / 8 * bits, din, dout,
                                                                 count, bits, levels, nlevels, s);
                             });
}

int fewbit_piecewise_forward_host(int func, int dtype, const void *x_host, void *y_host, uint8_t *state,
                                  int64_t n, double p0, double p1, int64_t chunk_elems) {
    if (func < 0 || func >= FEWBIT_NUM_PIECEWISE) return FEWBIT_EFUNC;
    if (int st = check_common(dtype, x_host, y_host, state, n)) return st;
    return run_host_pipeline(dtype, x_host, y_host, n, chunk_elems,
                             [&](void *din, void *dout, int64_t first, int64_t count, cudaStream_t s) {
                                 return fewbit_piecewise_forward(func, dtype, din, dout, state + first / 8,
                                                                 count, p0, p1, s);
                             });
}

int fewbit_piecewise_backward_host(int func, int dtype, const uint8_t *state, const void *gout_host,
                                   void *gin_host, int64_t n, double p0, int64_t chunk_elems) {
    if (func < 0 || func >= FEWBIT_NUM_PIECEWISE) return FEWBIT_EFUNC;
    if (int st = check_common(dtype, gout_host, gin_host, state, n)) return st;
    return run_host_pipeline(dtype, gout_host, gin_host, n, chunk_elems,
                             [&](void *din, void *dout, int64_t first, int64_t count, cudaStream_t s) {
                                 return fewbit_piecewise_backward(func, dtype, state + first / 8, din, dout,
                                                                  count, p0, s);
                             });
}

int64_t fewbit_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
