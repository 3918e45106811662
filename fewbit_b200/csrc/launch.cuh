// fewbit_b200 -- host-side launch helpers (grid sizing, aligned/ragged split).
#pragma once

#include <algorithm>
#include <atomic>
#include <type_traits>
#include <cstdlib>

#include "ops.cuh"

// Tile configuration per kernel class, from the sweeps in profiles/r01_sweep_{a,b,final}.txt
// (B200, GB/s of algorithmic bytes; torch's copy kernel reaches 6520 on the same box):
//   light kernels (1-bit masks, every backward): two LDG.128 in flight per lane and no register
//     cap -> 6.2-6.3 TB/s; four loads or a tighter register budget cost 5-7 %.
//   math-heavy forward kernels (continuous activations): U = 4 subtiles per warp tile with a
//     64 (bf16) / 80 (fp32) register budget.
// Overridable per build for further sweeps.
#ifndef FEWBIT_U_LIGHT_F32
#define FEWBIT_U_LIGHT_F32 1
#endif
#ifndef FEWBIT_U_LIGHT_BF16
#define FEWBIT_U_LIGHT_BF16 2
#endif
#ifndef FEWBIT_U_HEAVY
#define FEWBIT_U_HEAVY 4
#endif
#ifndef FEWBIT_TILE_HINTS
#define FEWBIT_TILE_HINTS 1   // 0: one tile shape for every fp32 forward kernel (A/B builds)
#endif
#ifndef FEWBIT_MAX_CTAS_PER_SM
#define FEWBIT_MAX_CTAS_PER_SM 4
#endif
#ifndef FEWBIT_MINB_LIGHT
#define FEWBIT_MINB_LIGHT 1
#endif
#ifndef FEWBIT_MINB_HEAVY_F32
#define FEWBIT_MINB_HEAVY_F32 3
#endif
#ifndef FEWBIT_MINB_HEAVY_BF16
#define FEWBIT_MINB_HEAVY_BF16 4
#endif

namespace fewbit {

// Arguments of the generic entry points, type-erased (see include/fewbit_b200.h).
struct ForwardArgs {
    int dtype;
    const void *x;
    void *y;
    uint8_t *state;
    int64_t n;
    int bits;
    const void *table;  // bounds
    int ntable;
    double p0, p1;
    cudaStream_t stream;
};
struct BackwardArgs {
    int dtype;
    const uint8_t *state;
    const void *gout;
    void *gin;
    int64_t n;
    int bits;
    const void *table;  // levels
    int ntable;
    double p0;
    cudaStream_t stream;
};

void note_launch();  // api.cu: bumps fewbit_launch_count()
int sm_count();      // api.cu: SM count of the current device (cached per device)

template <class Op, typename T> struct TileConfig {
    static constexpr bool kHeavy = Op::kHeavy;
    // Resident CTAs per SM the persistent grid is sized for.  The light fp32 kernels fit five, but
    // four (32 warps, 32 KB of loads in flight per SM) is measurably the better operating point
    // of the memory system: backward 88 % -> 94-95 % of the HBM peak on a 200 MB tensor, forward
    // 86 % -> 91-93 % (profiles/r01_function_sweep_3bit.md).  The bf16 kernels sit at four anyway.
    static constexpr int kMaxCtasPerSm = FEWBIT_MAX_CTAS_PER_SM;
    template <class O, typename = void> struct Hint {   // ops without a per-function hint
        static constexpr int kSubtiles = FEWBIT_U_HEAVY, kMinBlocks = FEWBIT_MINB_HEAVY_F32;
    };
    template <class O> struct Hint<O, std::enable_if_t<(O::kSubtilesF32 > 0)>> {
        static constexpr int kSubtiles = FEWBIT_TILE_HINTS ? O::kSubtilesF32 : FEWBIT_U_HEAVY;
        static constexpr int kMinBlocks = FEWBIT_TILE_HINTS ? O::kMinBlocksF32 : FEWBIT_MINB_HEAVY_F32;
    };
    static constexpr int kSubtiles =
        kHeavy ? (sizeof(T) == 4 && !wants_stream_f32<Op>::value ? Hint<Op>::kSubtiles : FEWBIT_U_HEAVY)
               : (sizeof(T) == 2 ? FEWBIT_U_LIGHT_BF16 : FEWBIT_U_LIGHT_F32);
    static constexpr int kMinBlocks =
        kHeavy ? (sizeof(T) == 2 ? FEWBIT_MINB_HEAVY_BF16 : Hint<Op>::kMinBlocks) : FEWBIT_MINB_LIGHT;
};

// Resident CTAs per SM for `kernel` (occupancy API, cached per instantiation), overridable
// with FEWBIT_B200_CTAS_PER_SM for tuning runs.
template <auto kernel> int resident_ctas(int most, int dynamic_smem = 0) {
    // the answer depends on the kernel and the architecture only (sm_100a everywhere): caching it
    // per instantiation, not per device, is enough; a racing first call computes the same value
    static std::atomic<int> cached{0};
    if (cached.load(std::memory_order_relaxed) == 0) {
        int blocks = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, kThreads, dynamic_smem) !=
                cudaSuccess ||
            blocks < 1)
            blocks = 1;
        blocks = std::min(blocks, most);
        if (const char *env = std::getenv("FEWBIT_B200_CTAS_PER_SM")) {
            int v = std::atoi(env);
            if (v > 0) blocks = std::min(blocks, v);
        }
        cached.store(blocks, std::memory_order_relaxed);
    }
    return cached.load(std::memory_order_relaxed);
}

template <typename T> bool vector_aligned(const void *a, const void *b, const void *c) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
             reinterpret_cast<uintptr_t>(c)) & 15u) == 0;
}

template <class Op, typename T>
cudaError_t launch_forward(const T *x, T *y, uint8_t *state, int64_t n, const Op &op,
                           cudaStream_t stream) {
    constexpr int U = TileConfig<Op, T>::kSubtiles;
    constexpr int64_t kTile = (int64_t)U * kSubtile;
    if (n <= 0) return cudaSuccess;
    const int64_t ntiles = vector_aligned<T>(x, y, state) ? n / kTile : 0;
    if (ntiles > 0) {
        constexpr auto kernel = forward_tiles_kernel<Op, T, U, TileConfig<Op, T>::kMinBlocks>;
        constexpr int kRing = ring_bytes<Op, T, U>();       // dynamic shared memory: the cp.async input ring
        if constexpr (kRing > 0) {
            // static + dynamic shared memory may pass the 48 KB default: opt in once per device
            static std::atomic<unsigned long long> configured{0};
            int device = 0;
            cudaGetDevice(&device);
            const unsigned long long bit = 1ull << (device & 63);
            if (!(configured.load(std::memory_order_relaxed) & bit)) {
                cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRing);
                if (e != cudaSuccess) return e;
                configured.fetch_or(bit, std::memory_order_relaxed);
            }
        }
        const int64_t want = (ntiles + kWarps - 1) / kWarps;
        const int64_t cap = (int64_t)sm_count() * resident_ctas<kernel>(TileConfig<Op, T>::kMaxCtasPerSm, kRing);
        kernel<<<(unsigned)std::min(want, cap), kThreads, kRing, stream>>>(x, y, state, ntiles, op);
        note_launch();
    }
    const int64_t first = ntiles * kTile;
    if (first < n) {
        auto kernel = forward_ragged_kernel<Op, T>;
        const int64_t octets = (n - first + 7) / 8;
        const int64_t want = (octets + kThreads - 1) / kThreads;
        const int64_t cap = (int64_t)sm_count() * 8;
        kernel<<<(unsigned)std::min(want, cap), kThreads, 0, stream>>>(x, y, state, first, n, op);
        note_launch();
    }
    return cudaGetLastError();
}

template <class Op, typename T>
cudaError_t launch_backward(const uint8_t *state, const T *gout, T *gin, int64_t n, const Op &op,
                            cudaStream_t stream) {
    constexpr int U = TileConfig<Op, T>::kSubtiles;
    constexpr int64_t kTile = (int64_t)U * kSubtile;
    if (n <= 0) return cudaSuccess;
    const int64_t ntiles = vector_aligned<T>(gout, gin, state) ? n / kTile : 0;
    if (ntiles > 0) {
        constexpr auto kernel = backward_tiles_kernel<Op, T, U, TileConfig<Op, T>::kMinBlocks>;
        const int64_t want = (ntiles + kWarps - 1) / kWarps;
        const int64_t cap = (int64_t)sm_count() * resident_ctas<kernel>(TileConfig<Op, T>::kMaxCtasPerSm);
        kernel<<<(unsigned)std::min(want, cap), kThreads, 0, stream>>>(state, gout, gin, ntiles,
                                                                       op);
        note_launch();
    }
    const int64_t first = ntiles * kTile;
    if (first < n) {
        auto kernel = backward_ragged_kernel<Op, T>;
        const int64_t octets = (n - first + 7) / 8;
        const int64_t want = (octets + kThreads - 1) / kThreads;
        const int64_t cap = (int64_t)sm_count() * 8;
        kernel<<<(unsigned)std::min(want, cap), kThreads, 0, stream>>>(state, gout, gin, first, n,
                                                                       op);
        note_launch();
    }
    return cudaGetLastError();
}

// Expand `body(B)` for the run-time bit width.
#define FEWBIT_DISPATCH_BITS(bits, BODY) \
    switch (bits) {                      \
        case 1: { constexpr int B = 1; BODY; } break; \
        case 2: { constexpr int B = 2; BODY; } break; \
        case 3: { constexpr int B = 3; BODY; } break; \
        case 4: { constexpr int B = 4; BODY; } break; \
        case 5: { constexpr int B = 5; BODY; } break; \
        case 6: { constexpr int B = 6; BODY; } break; \
        case 7: { constexpr int B = 7; BODY; } break; \
        case 8: { constexpr int B = 8; BODY; } break; \
        default: break;                  \
    }

}  // namespace fewbit
