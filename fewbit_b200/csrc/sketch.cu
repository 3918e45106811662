// fewbit_b200 -- RandomizedLinear's projection  out = scale * S X  on tcgen05 tensor cores.
//
//   X   : [N tokens, D features] bf16, row-major (TMA source; never copied or transposed)
//   S   : [P, N] random sketch, N(0,1) or +-1/2 entries, NEVER materialised: every element is
//         a pure function of (seed, offset, p, n) -- Philox4x32-10 keyed by the seed; one call
//         with counter (n / 8, p, offset) gives eight normals (Box-Muller on 16-bit uniforms),
//         one call with counter (n / 128, p, offset) gives 128 signs -- so forward and backward
//         regenerate the same S.
//   out : [P, D] fp32
// Replaces `proj = randn(P, N); proj_input = (proj @ input_view) / P` and `proj @ grad_output`
// of the reference (fewbit/functional/linear.py:133-137, 196-199), which writes S (214 MB at
// RoBERTa shapes) to HBM twice per layer and multiplies in fp32 on the CUDA cores.
//
// Mapping onto the MMA.  The contraction runs over tokens, so with X as it lies in memory the
// feature axis is the contiguous one: X^T is an "MN-major" operand.  We therefore compute
// out^T tiles:  D[d, p] += A[d, n] * B[p, n]   with
//   A = X^T  : M = 128 features per MMA, MN-major, 128B swizzle, loaded by TMA as 64-token x
//              64-feature boxes (8 KB each, exactly the canonical MN-major SW128 atom stack);
//   B = S    : N = BN sketch rows (<= 160, multiple of 16), K-major, 128B swizzle, written to
//              shared memory by the generator warps;
//   D        : fp32 in TMEM, 128 lanes (features) x BN columns per 128-feature block; a CTA owns
//              up to 3 blocks (384 features -> 480 of the 512 TMEM columns), so one generated S
//              tile feeds three MMAs and X is re-read from L2 only P/BN times.
// The TMEM lane = feature orientation also makes the epilogue store coalesced: for a fixed
// sketch row the 32 lanes of a warp hold 32 consecutive features.
//
// CTA = 16 warps: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation), warps 2-15
// generate S (the generation is the most expensive part: ~17 instructions and 2 MUFU ops per
// normal, so it gets 14 of the 16 warps and several independent Philox chains per thread);
// warps 4-7 then run the epilogue (tcgen05.ld -> scale -> global).  Three-stage mbarrier
// pipeline: full_x (TMA bytes), full_s (generator warps), empty (tcgen05.commit).
// Grid = (ceil(P / BN), ceil(D / 384), split_k); split-K partials are reduced by a tiny kernel.
//
// Sharing operands inside a thread-block cluster (Cx, Cy, 1), Cx, Cy in {1, 2}:
//  * along y (feature tiles): generating S costs about twice the MMA time, and CTAs that differ
//    only in their feature tile need the very same S tile.  CTA ry generates rows
//    [ry BN/Cy, (ry+1) BN/Cy) of each stage into its own shared memory and pushes that block to
//    its y-peer with cp.async.bulk.shared::cluster (async proxy; completes transaction bytes on
//    the PEER's full_s barrier, so the tensor core sees the data without a generic-proxy
//    hand-over).
//  * along x (sketch-row tiles): CTAs that differ only in their row tile stream the very same X
//    tiles, and re-reading X from L2 once per row tile is what bounds the kernel when S is cheap
//    (Rademacher).  Each CTA loads every Cx-th TMA box and multicasts it to its x-peers.
// A stage may be overwritten only when every CTA of the cluster has consumed it: tcgen05.commit
// multicasts its arrival to the `empty` barrier of all CTAs of the cluster.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "../../include/fewbit_b200.h"

namespace fewbit {

void note_launch();
int sm_count();

namespace sketch {

constexpr int kThreads = 512;
constexpr int kStages = 3;
constexpr int kBlockK = 64;            // tokens per stage
constexpr int kFeaturesPerCta = 384;   // 3 MMA M-blocks of 128
constexpr int kMaxRows = 160;          // BN: sketch rows per CTA (TMEM: 3 * 160 <= 512 columns)
constexpr int kBoxBytes = 64 * 64 * 2;                        // one TMA box: 64 tokens x 64 features
constexpr int kXStageBytes = (kFeaturesPerCta / 64) * kBoxBytes;   // 49152
constexpr int kSStageBytes = kMaxRows * 128;                  // 20480: BN rows x 64 bf16
constexpr int kStageBytes = kXStageBytes + kSStageBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /* alignment */ + 256 /* barriers */;
constexpr int kGeneratorWarps = 14;
constexpr int kGeneratorThreads = kGeneratorWarps * 32;
constexpr int kTmemColumns = 512;

// ---------------------------------------------------------------------------- PTX ----

__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst, const CUtensorMap *map, int c0, int c1,
                                                      uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_cta_y() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctaid.y;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {   // shared::cta -> shared::cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes,
                                                  uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            dst_cluster),
        "r"(src_cta), "r"(bytes), "r"(bar_cluster)
        : "memory");
}
__device__ __forceinline__ void umma_commit_cluster(uint32_t bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_load16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100): start address, leading
// and stride byte offsets in 16-byte units, version 1, layout SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------------- random ----

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ uint4 operator()(uint4 c) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int round = 0; round < 10; ++round) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
            c = make_uint4(hi1 ^ c.y ^ a, lo1, hi0 ^ c.w ^ b, lo0);
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        return c;
    }
};

__device__ __forceinline__ float lg2_approx(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// (h + 1/2) / 65536 in (0, 1) from 16 random bits: mantissa trick, no integer->float convert.
__device__ __forceinline__ float unit16(uint32_t h) {
    return __uint_as_float(0x3f800000u | (h << 7)) - (1.0f - 7.62939453125e-6f);
}
// Two normals from one 32-bit word: Box-Muller with a 16-bit radius and a 16-bit angle.
__device__ __forceinline__ uint32_t normal_pair(uint32_t word) {
    const float radius = sqrt_approx(-1.3862943611198906f * lg2_approx(unit16(word & 0xffffu)));   // sqrt(-2 ln u)
    const float angle = 6.283185307179586f * unit16(word >> 16);
    return pack_bf16(radius * __cosf(angle), radius * __sinf(angle));
}
// kind 0: entries S[p][8o .. 8o+7] (eight normals) as four packed bf16x2 words.
__device__ __forceinline__ uint4 normal_octet(const Philox &rng, uint32_t o, uint32_t p, uint32_t off_lo,
                                              uint32_t off_hi) {
    const uint4 r = rng(make_uint4(o, p, off_lo, off_hi));
    return make_uint4(normal_pair(r.x), normal_pair(r.y), normal_pair(r.z), normal_pair(r.w));
}
// kind 1: 128 signs S[p][128c .. 128c+127]; bit b of word w is entry 32w + b.
__device__ __forceinline__ uint4 sign_block(const Philox &rng, uint32_t c, uint32_t p, uint32_t off_lo,
                                            uint32_t off_hi) {
    return rng(make_uint4(c, p, off_lo, off_hi));
}
// Two consecutive sign bits -> packed bf16x2 of +-1/2 (0x3f00 = 0.5, sign bit set for bit = 0).
__device__ __forceinline__ uint32_t sign_pair(uint32_t bits) {
    return 0x3f003f00u | ((~bits & 1u) << 15) | ((~bits & 2u) << 30);
}

// --------------------------------------------------------------------------- kernel ----

struct Params {
    float *out;          // [P, D] (split_k == 1) or partials [split_k, P, D]
    int64_t tokens;      // N
    int features;        // D
    int rows;            // P
    int block_rows;      // BN
    int kblocks_per_split;
    int split_k;
    float scale;         // applied here only when split_k == 1
    uint32_t seed_lo, seed_hi, off_lo, off_hi;
    int kind;
    int cluster_x;       // CTAs along grid.x that share X tiles (TMA multicast)
    int cluster_y;       // CTAs along grid.y that share one generated S tile
    int debug;           // timing experiments only (results are garbage): 1 = skip generating S,
                         // 2 = skip loading X, 4 = skip issuing MMAs
};

__global__ void __launch_bounds__(kThreads, 1)
sketch_kernel(const __grid_constant__ CUtensorMap x_map, const Params prm) {
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled tiles need 1024-byte alignment.
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p0 = blockIdx.x * prm.block_rows;
    const int d0 = blockIdx.y * kFeaturesPerCta;
    const int bn = prm.block_rows;
    const int nblocks = min(3, (prm.features - d0 + 127) / 128);   // 128-feature MMA blocks
    const int nboxes = min(6, (prm.features - d0 + 63) / 64);
    const int64_t total_kb = (prm.tokens + kBlockK - 1) / kBlockK;
    const int64_t kb_begin = (int64_t)blockIdx.z * prm.kblocks_per_split;
    const int64_t kb_end = min(total_kb, kb_begin + prm.kblocks_per_split);
    const int iters = (int)max((int64_t)0, kb_end - kb_begin);

    auto full_x = [&](int s) { return smem_addr(bars + s); };
    auto full_s = [&](int s) { return smem_addr(bars + kStages + s); };
    auto empty = [&](int s) { return smem_addr(bars + 2 * kStages + s); };
    const uint32_t accum_full = smem_addr(bars + 3 * kStages);
    auto x_stage = [&](int s) { return smem_addr(smem + s * kStageBytes); };
    auto s_stage = [&](int s) { return smem_addr(smem + s * kStageBytes + kXStageBytes); };

    const int cx = prm.cluster_x, cy = prm.cluster_y, cluster = cx * cy;
    const uint32_t rx = cx > 1 ? cluster_cta_x() : 0, ry = cy > 1 ? cluster_cta_y() : 0;   // rank = rx + ry * cx
    const int my_rows = bn / cy;                            // rows of each S tile this CTA generates
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_x(s), 1);
            mbar_init(full_s(s), kGeneratorWarps + (cy > 1 ? 1 : 0));   // + the expect_tx arrival
            mbar_init(empty(s), cluster);                               // one commit per CTA
        }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM allocation is warp-wide; the same warp frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_addr(tmem_slot)),
                     "r"(kTmemColumns));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (cluster > 1) cluster_sync();          // peers' barriers exist before anyone signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer ----
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % kStages;
                mbar_wait(empty(s), ((it / kStages) & 1) ^ 1);
                if (prm.debug & 2) { mbar_arrive(full_x(s)); continue; }
                mbar_expect_tx(full_x(s), nboxes * kBoxBytes);                         // all boxes, whoever loads them
                if (cy > 1) mbar_expect_tx(full_s(s), (cy - 1) * my_rows * 128);       // the y-peer's block
                const int token = (int)((kb_begin + it) * kBlockK);
                if (cx > 1) {
                    const uint16_t mask = (uint16_t)(((1u << cx) - 1u) << (ry * cx));  // my row of the cluster
                    for (int b = (int)rx; b < nboxes; b += cx)
                        tma_load_2d_multicast(x_stage(s) + b * kBoxBytes, &x_map, d0 + 64 * b, token, full_x(s), mask);
                } else {
                    for (int b = 0; b < nboxes; ++b)
                        tma_load_2d(x_stage(s) + b * kBoxBytes, &x_map, d0 + 64 * b, token, full_x(s));
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer ----
        if (lane == 0) {
            // cute::UMMA::InstrDescriptor: D = f32, A = B = bf16, A MN-major, B K-major, N, M = 128.
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) |
                                   ((uint32_t)(bn >> 3) << 17) | ((128u >> 4) << 24);
            for (int it = 0; it < iters; ++it) {
                const int s = it % kStages;
                const uint32_t parity = (it / kStages) & 1;
                mbar_wait(full_x(s), parity);
                mbar_wait(full_s(s), parity);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < ((prm.debug & 4) ? 0 : kBlockK / 16); ++k) {
                    // B: rows of 128 B (64 tokens), 8-row groups 1024 B apart; +32 B per 16 tokens.
                    const uint64_t desc_b = smem_desc(s_stage(s) + 32 * k, 16, 1024);
                    for (int m = 0; m < nblocks; ++m) {
                        // A: 64-feature groups 8192 B apart (LBO), 8-token groups 1024 B apart
                        // (SBO); +2048 B per 16 tokens; 128 features = two boxes.
                        const uint64_t desc_a = smem_desc(x_stage(s) + m * 2 * kBoxBytes + 2048 * k, kBoxBytes, 1024);
                        umma_bf16(tmem_base + m * kMaxRows, desc_a, desc_b, idesc, (it | k) != 0);
                    }
                }
                // smem slot reusable once these MMAs have read it -- in every CTA of the cluster
                if (cluster > 1) umma_commit_cluster(empty(s), (uint16_t)((1u << cluster) - 1));
                else umma_commit(empty(s));
            }
            umma_commit(accum_full);          // accumulators complete
        }
    } else {
        // -------------------------------------------------------------- generators ----
        const Philox rng{prm.seed_lo, prm.seed_hi};
        const int gt = threadIdx.x - 64;                       // 0 .. 447
        for (int it = 0; it < iters; ++it) {
            const int s = it % kStages;
            mbar_wait(empty(s), ((it / kStages) & 1) ^ 1);
            uint8_t *tile = smem + s * kStageBytes + kXStageBytes;
            const int64_t kb = kb_begin + it;
            const int row0 = (int)ry * my_rows;                // this CTA's block of the tile
            if (prm.debug & 1) {
            } else if (prm.kind == 0) {
                // One Philox call = 8 normals = one 16-byte chunk of a 128-byte K-major row.
                // K-major SW128: chunk index XOR (row mod 8).  Up to 3 chunks per thread
                // (160 rows x 8 chunks over 448 threads), independent chains interleave.
                const int chunks = my_rows * 8;
#pragma unroll
                for (int j = 0; j < (kMaxRows * 8 + kGeneratorThreads - 1) / kGeneratorThreads; ++j) {
                    const int i = gt + j * kGeneratorThreads;
                    if (i < chunks) {
                        const int row = row0 + (i >> 3), o = i & 7;
                        const uint4 v = normal_octet(rng, (uint32_t)(kb * 8 + o), (uint32_t)(p0 + row),
                                                     prm.off_lo, prm.off_hi);
                        *reinterpret_cast<uint4 *>(tile + (row >> 3) * 1024 + (row & 7) * 128 + ((o ^ (row & 7)) << 4)) = v;
                    }
                }
            } else {
                // One Philox call = 128 signs; a 64-token stage uses half of it (two words), and
                // each task expands one word = 32 tokens = four 16-byte chunks of a row, so that
                // 2 x rows tasks keep most generator threads busy (the call is recomputed by the
                // two tasks of a row: cheaper than leaving 3/4 of the threads idle).
                for (int task = gt; task < 2 * my_rows; task += kGeneratorThreads) {
                    const int row = row0 + (task >> 1), half = task & 1;
                    const uint4 w = sign_block(rng, (uint32_t)(kb >> 1), (uint32_t)(p0 + row), prm.off_lo, prm.off_hi);
                    const uint32_t word = (kb & 1) ? (half ? w.w : w.z) : (half ? w.y : w.x);
                    uint8_t *base = tile + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t bits = word >> (c * 8);
                        const int o = half * 4 + c;
                        *reinterpret_cast<uint4 *>(base + ((o ^ (row & 7)) << 4)) =
                            make_uint4(sign_pair(bits), sign_pair(bits >> 2), sign_pair(bits >> 4), sign_pair(bits >> 6));
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic -> async proxy
            if (cy > 1) {
                // all generator threads have written (and fenced) this CTA's block: push it to the y-peers
                asm volatile("bar.sync 1, %0;" ::"n"(kGeneratorThreads) : "memory");
                if (gt == 0) {
                    const uint32_t block = smem_addr(tile) + row0 * 128;
                    for (uint32_t y = 0; y < (uint32_t)cy; ++y)
                        if (y != ry)
                            bulk_copy_to_peer(map_to_cta(block, rx + y * cx), block, my_rows * 128,
                                              map_to_cta(full_s(s), rx + y * cx));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(full_s(s));
        }
        // ---------------------------------------------------------------- epilogue ----
        if (warp >= 4 && warp < 8) {
            mbar_wait(accum_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int quarter = warp & 3;                       // TMEM lanes [32q, 32q + 32)
            float *out = prm.out + (prm.split_k > 1 ? (int64_t)blockIdx.z * prm.rows * prm.features : 0);
            const float scale = prm.split_k > 1 ? 1.0f : prm.scale;
            for (int m = 0; m < nblocks; ++m) {
                const int d = d0 + m * 128 + quarter * 32 + lane;
                for (int c = 0; c < bn; c += 16) {
                    uint32_t v[16];
                    tmem_load16(tmem_base + ((uint32_t)(quarter * 32) << 16) + m * kMaxRows + c, v);
                    if (iters == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = 0;
                    }
                    if (d < prm.features) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int p = p0 + c + j;
                            if (p < prm.rows) out[(int64_t)p * prm.features + d] = __uint_as_float(v[j]) * scale;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (cluster > 1) cluster_sync();          // nobody leaves while peers may still signal it
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(kTmemColumns));
    }
}

__global__ void reduce_splits_kernel(const float *partials, float *out, int64_t count, int splits,
                                     float scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        float acc = 0.0f;
        for (int s = 0; s < splits; ++s) acc += partials[(int64_t)s * count + i];
        out[i] = acc * scale;
    }
}

// S itself, for tests and diagnostics only (the product never materialises it).
__global__ void sketch_matrix_kernel(__nv_bfloat16 *s, int rows, int64_t cols, Params prm) {
    const Philox rng{prm.seed_lo, prm.seed_hi};
    const int64_t octets_per_row = (cols + 7) / 8;
    uint16_t *dst = reinterpret_cast<uint16_t *>(s);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * octets_per_row;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / octets_per_row, o = i % octets_per_row;
        uint32_t w[4];
        if (prm.kind == 0) {
            const uint4 v = normal_octet(rng, (uint32_t)o, (uint32_t)p, prm.off_lo, prm.off_hi);
            w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
        } else {
            const uint4 r = sign_block(rng, (uint32_t)(o >> 4), (uint32_t)p, prm.off_lo, prm.off_hi);
            const uint32_t words[4] = {r.x, r.y, r.z, r.w};
            const uint32_t bits = words[(o >> 2) & 3] >> ((o & 3) * 8);
            w[0] = sign_pair(bits), w[1] = sign_pair(bits >> 2), w[2] = sign_pair(bits >> 4), w[3] = sign_pair(bits >> 6);
        }
        for (int j = 0; j < 8; ++j)
            if (8 * o + j < cols) dst[p * cols + 8 * o + j] = (uint16_t)(w[j >> 1] >> ((j & 1) * 16));
    }
}

// ----------------------------------------------------------------------------- host ----

using EncodeTiled = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);

static EncodeTiled encode_tiled() {
    static EncodeTiled fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &status) == cudaSuccess &&
            status == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiled>(ptr);
    }
    return fn;
}

// Pick the cluster (Cx row tiles sharing X by TMA multicast, Cy feature tiles sharing one
// generated S tile), BN (multiple of 16 and of 8 Cy, <= 160) and split_k: minimise
// waves * (time of one CTA).  Measured on B200 (profiles/r01_sketch_kernel.md): clusters of 8
// along y lose to independent CTAs at D = 3072 (8-CTA placement leaves SMs idle and the lock-step
// `empty` barrier couples eight pipelines), so both axes are capped at 2.
static void plan(int rows, int features, int64_t tokens, int sms, int &bn, int &split_k, int &cx, int &cy) {
    const int dtiles = (features + kFeaturesPerCta - 1) / kFeaturesPerCta;
    const int64_t kblocks = std::max<int64_t>(1, (tokens + kBlockK - 1) / kBlockK);
    // Default from the A/B runs in profiles/r01_sketch_kernel.md: S sharing between the two feature
    // tiles of D = 768 helps (133 -> 123 us); X multicast does not (159 us), L2 is not the limiter.
    cy = dtiles == 2 ? 2 : 1;
    cx = 1;
    if (const char *env = std::getenv("FEWBIT_B200_SKETCH_CLUSTER")) {   // A/B runs: "<cx><cy>", e.g. 11, 21, 12, 22
        const int v = std::atoi(env);
        if (v / 10 >= 1 && v / 10 <= 2) cx = rows > 160 ? v / 10 : 1;
        if (v % 10 >= 1 && v % 10 <= 2) cy = std::min(cy, v % 10);
    }
    double best = 1e300;
    bn = 64, split_k = 1;
    for (int cand = 160; cand >= 64; cand -= 16) {
        if ((cand / 8) % cy != 0) continue;
        const int ptiles = ((rows + cand - 1) / cand + cx - 1) / cx * cx;
        for (int sk = 1; sk <= 8 && sk <= kblocks; ++sk) {
            const int64_t ctas = (int64_t)ptiles * dtiles * sk;
            const int64_t waves = (ctas + sms - 1) / sms;
            // per 64-token block: S generation ~11 cycles per generated row, MMA 6 cycles per row,
            // and the 48 KB X tile needs ~1100 cycles to arrive from L2 (half with multicast)
            const double block = std::max({cand * 11.0 / cy, cand * 6.0, 1100.0 / cx});
            const double per_cta = (double)((kblocks + sk - 1) / sk) * block + 8000.0;
            const double cost = (double)waves * per_cta * (1.0 + 0.01 * (sk - 1));
            if (cost < best) best = cost, bn = cand, split_k = sk;
        }
    }
}

}  // namespace sketch
}  // namespace fewbit

using namespace fewbit;
using namespace fewbit::sketch;

extern "C" {

size_t fewbit_sketch_workspace_bytes(int64_t tokens, int features, int rows) {
    int bn, split_k, cx, cy;
    plan(rows, features, tokens, sm_count(), bn, split_k, cx, cy);
    return split_k > 1 ? (size_t)split_k * rows * features * sizeof(float) : 0;
}

int fewbit_sketch_forward(const void *x, float *out, void *workspace, int64_t tokens, int features,
                          int rows, int kind, float scale, uint64_t seed, uint64_t offset, void *stream) {
    if (tokens < 0 || features <= 0 || rows <= 0 || (kind != 0 && kind != 1)) return FEWBIT_EINVAL;
    if (!x || !out) return FEWBIT_EINVAL;
    if (features % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15)) return FEWBIT_EALIGN;  // TMA strides
    EncodeTiled encode = encode_tiled();
    if (!encode) return (int)cudaErrorNotSupported;
    cudaStream_t s = (cudaStream_t)stream;
    int bn, split_k, cx, cy;
    plan(rows, features, tokens, sm_count(), bn, split_k, cx, cy);
    if (split_k > 1 && !workspace) return FEWBIT_EINVAL;

    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)features, (cuuint64_t)std::max<int64_t>(tokens, 1)};
    const cuuint64_t strides[1] = {(cuuint64_t)features * 2};
    const cuuint32_t box[2] = {64, 64}, elem[2] = {1, 1};
    if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(x), dims, strides, box, elem,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return (int)cudaErrorInvalidValue;

    const int64_t kblocks = (tokens + kBlockK - 1) / kBlockK;
    Params prm;
    prm.out = split_k > 1 ? static_cast<float *>(workspace) : out;
    prm.tokens = tokens, prm.features = features, prm.rows = rows, prm.block_rows = bn;
    prm.kblocks_per_split = (int)((kblocks + split_k - 1) / split_k);
    prm.split_k = split_k, prm.scale = scale, prm.kind = kind, prm.cluster_x = cx, prm.cluster_y = cy;
    prm.debug = 0;
    if (const char *env = std::getenv("FEWBIT_B200_SKETCH_DEBUG")) prm.debug = std::atoi(env);
    prm.seed_lo = (uint32_t)seed, prm.seed_hi = (uint32_t)(seed >> 32);
    prm.off_lo = (uint32_t)offset, prm.off_hi = (uint32_t)(offset >> 32);

    // function attributes live in the device's context: set once per device
    static bool configured[64] = {};
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64) device = 0;
    if (!configured[device]) {
        cudaError_t e = cudaFuncSetAttribute(sketch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        configured[device] = true;
    }
    cudaLaunchConfig_t cfg{};
    // grid.x rounded up to whole clusters: surplus CTAs compute rows >= P, which are never stored
    cfg.gridDim = dim3(((rows + bn - 1) / bn + cx - 1) / cx * cx, (features + kFeaturesPerCta - 1) / kFeaturesPerCta, split_k);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cx, attr[0].val.clusterDim.y = cy, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    cudaError_t launched = cudaLaunchKernelEx(&cfg, sketch_kernel, map, prm);
    if (launched != cudaSuccess) return (int)launched;
    note_launch();
    if (split_k > 1) {
        const int64_t count = (int64_t)rows * features;
        reduce_splits_kernel<<<(unsigned)std::min<int64_t>((count + 255) / 256, sm_count() * 8), 256, 0, s>>>(
            static_cast<const float *>(workspace), out, count, split_k, scale);
        note_launch();
    }
    return (int)cudaGetLastError();
}

int fewbit_sketch_matrix(void *s_bf16, int rows, int64_t cols, int kind, uint64_t seed, uint64_t offset,
                         void *stream) {
    if (!s_bf16 || rows <= 0 || cols <= 0 || (kind != 0 && kind != 1)) return FEWBIT_EINVAL;
    Params prm{};
    prm.kind = kind;
    prm.seed_lo = (uint32_t)seed, prm.seed_hi = (uint32_t)(seed >> 32);
    prm.off_lo = (uint32_t)offset, prm.off_hi = (uint32_t)(offset >> 32);
    const int64_t octets = rows * ((cols + 7) / 8);
    sketch_matrix_kernel<<<(unsigned)std::min<int64_t>((octets + 255) / 256, sm_count() * 16), 256, 0,
                           (cudaStream_t)stream>>>(static_cast<__nv_bfloat16 *>(s_bf16), rows, cols, prm);
    note_launch();
    return (int)cudaGetLastError();
}

}  // extern "C"
