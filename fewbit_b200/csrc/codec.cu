// fewbit_b200 -- stand-alone codec on int32 codes (exact-`bits` stream), one thread per octet.
// Replaces DeflateBlock / InflateBlock (reference fewbit/cuda/codec.cu:166-220); not on the
// training hot path (the fused kernels never materialise int32 codes).
#include "launch.cuh"

namespace fewbit {

__global__ void __launch_bounds__(kThreads) deflate_kernel(const int32_t *codes, uint8_t *state,
                                                          int64_t n, int bits) {
    const int64_t nbytes = (n * bits + 7) / 8, noctets = (n + 7) / 8;
    const uint32_t mask = (1u << bits) - 1u;
    for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < noctets;
         o += (int64_t)gridDim.x * kThreads) {
        uint64_t octet = 0;
        for (int j = 0; j < 8; ++j)
            if (8 * o + j < n) octet |= (uint64_t)((uint32_t)codes[8 * o + j] & mask) << (bits * j);
        for (int k = 0; k < bits; ++k)
            if (o * bits + k < nbytes) state[o * bits + k] = (uint8_t)(octet >> (8 * k));
    }
}

__global__ void __launch_bounds__(kThreads) inflate_kernel(const uint8_t *state, int32_t *codes,
                                                          int64_t n, int bits) {
    const int64_t nbytes = (n * bits + 7) / 8, noctets = (n + 7) / 8;
    const uint32_t mask = (1u << bits) - 1u;
    for (int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x; o < noctets;
         o += (int64_t)gridDim.x * kThreads) {
        uint64_t octet = 0;
        for (int k = 0; k < bits; ++k)
            if (o * bits + k < nbytes) octet |= (uint64_t)state[o * bits + k] << (8 * k);
        for (int j = 0; j < 8; ++j)
            if (8 * o + j < n) codes[8 * o + j] = (int32_t)((uint32_t)(octet >> (bits * j)) & mask);
    }
}

static unsigned octet_grid(int64_t n) {
    const int64_t want = ((n + 7) / 8 + kThreads - 1) / kThreads;
    return (unsigned)std::min<int64_t>(want, (int64_t)sm_count() * 8);
}

cudaError_t launch_deflate(const int32_t *codes, uint8_t *state, int64_t n, int bits,
                           cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    deflate_kernel<<<octet_grid(n), kThreads, 0, stream>>>(codes, state, n, bits);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_inflate(const uint8_t *state, int32_t *codes, int64_t n, int bits,
                           cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    inflate_kernel<<<octet_grid(n), kThreads, 0, stream>>>(state, codes, n, bits);
    note_launch();
    return cudaGetLastError();
}

}  // namespace fewbit
