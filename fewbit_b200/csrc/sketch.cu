// fewbit_b200 -- RandomizedLinear's projection  out = scale * S X  on tcgen05 tensor cores.
//
//   X   : [N tokens, D features] bf16, row-major (TMA source; never copied or transposed)
//   S   : [P, N] random sketch, N(0,1) or +-1/2 entries, NEVER materialised: every element is
//         a pure function of (seed, offset, p, n) -- Philox4x32-7 keyed by the seed; one call
//         with counter (n / 8, p, offset) gives eight normals (Box-Muller on 16-bit uniforms),
//         one call with counter (n / 128, p, offset) gives 128 signs -- so forward and backward
//         regenerate the same S.
//   out : [P, D] fp32 (fewbit_sketch_forward) or fp32 / bf16 with an optional extra row of column sums
//         (fewbit_sketch_project: one row of S is all ones -- the bias gradient of the backward pass)
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
// generate S; all 16 warps run the epilogue (tcgen05.ld -> scale -> global).  Two rings in shared
// memory, each with full/empty mbarriers (TMA bytes or generator warps / tcgen05.commit):
//   X ring: three stages of 64 tokens x 384 features (48 KB each);
//   S ring: slots of 128 tokens (two 64-token K-major tiles), as many as fit beside the X ring
//           (4 in pair mode, 2 otherwise).  The generators run up to that many slots ahead of the
//           MMAs, so neither their latency nor the hand-over of a slot between the CTAs of a pair
//           sits on the round trip of an X stage (profiles/r02_sketch_trace.md: with one ring the
//           MMA warp waited for S 60 % of the time while the generators were busy 75 % of it).
// Gaussian S: a generator thread owns the same (row, 16-byte chunk) positions of every slot and makes
// the Philox calls of the NEXT slot in the same straight-line block as the Box-Muller evaluations
// of this one, dealt out between the Philox rounds: multiply-xor rounds (ALU / FMA pipes) and
// lg2 / sqrt / sin / cos chains (XU pipe) overlap instead of taking turns.
// The TMA and MMA loops are walked by their whole warp with only the asynchronous instruction under
// an elect.sync predicate: under `if (lane == 0)` the compiler wraps every UTCHMMA in a uniform-
// register retry loop that costs ~150 cycles per MMA, twice the MMA itself (measured).
// Grid = (ceil(P / BN), ceil(D / 384), split_k); split-K partials are reduced by a tiny kernel.
//
// Two ways for CTAs that need the same S to share it (chosen by plan()):
//  * pair mode, whenever D is a multiple of 768: a 2 x 1 cluster is a CTA pair in the tcgen05 sense
//    (cta_group::2).  The two CTAs own the two 384-feature halves of a 768-feature slab and generate
//    HALF of every S slot each; one M = 256 MMA, issued by the leader, reads X^T from both shared
//    memories and each half of S once for both.  S is thus generated once per slab instead of once
//    per feature tile, nothing is copied, and the MMA's shared-memory traffic per SM drops by a
//    third.  The leader's full_x counts the TMA bytes of both CTAs (the peer's TMA signals the
//    leader's barrier, .cta_group::2 form); the peer's generators arrive on a local barrier and its
//    idle warp 1 forwards that to the leader with a cluster-scope release (~0.5 us, off the
//    generators' critical path).  Grid axes are swapped (feature tiles along x): the hardware pairs
//    CTAs that are adjacent along x of the cluster.
//  * along y (two feature tiles that are not a full pair) without cta_group::2, cluster (1, 2): CTA
//    ry generates rows [ry BN/2, (ry+1) BN/2) of each slot into its own shared memory and pushes
//    that block to its y-peer with cp.async.bulk.shared::cluster (async proxy; completes transaction
//    bytes on the PEER's full_s barrier, so the tensor core sees the data without a generic-proxy
//    hand-over).
// Tried and dropped (profiles/r01_sketch_kernel.md, r02_sketch_unit_ring.txt): sharing X between row
// tiles by TMA multicast, clusters of 8 along y, an X ring in 16 KB units with one commit each, L2
// prefetch ahead of the TMA, the leader loading the peer's X boxes, more generator warps, split-K summed inside a
// 2 x 1 x 3 cluster through distributed shared memory (profiles/r02_sketch_trace.md).
// A stage or slot may be overwritten only when every CTA of the cluster has consumed it:
// tcgen05.commit multicasts its arrival to the `empty` barriers of all CTAs of the cluster.
// FEWBIT_B200_SKETCH_TRACE=1 prints one CTA's timeline per call (benchmarks/sketch_trace.py).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "../../include/fewbit_b200.h"

namespace fewbit {

void note_launch();
int sm_count();

namespace sketch {

#ifndef FEWBIT_SKETCH_THREADS
#define FEWBIT_SKETCH_THREADS 512
#endif
constexpr int kThreads = FEWBIT_SKETCH_THREADS;     // warp 0: TMA, warp 1: MMA, the rest generate S
constexpr int kBlockK = 64;            // tokens per stage (four K = 16 MMAs per feature block)
constexpr int kFeaturesPerCta = 384;   // 3 MMA M-blocks of 128
constexpr int kMaxRows = 160;          // BN: sketch rows per CTA (TMEM: 3 * 160 <= 512 columns)
constexpr int kBoxBytes = 64 * 64 * 2;                        // one TMA box: 64 tokens x 64 features
constexpr int kStages = 3;                                    // X ring: stages of 64 tokens x 384 features
constexpr int kXStageBytes = (kFeaturesPerCta / 64) * kBoxBytes;   // 49152
constexpr int kMaxSlots = 4;                                  // S ring: slots of 128 tokens, count chosen at launch
constexpr int kBarrierBytes = 512;
constexpr int kSmemLimit = 232448;                            // 227 KB per CTA
constexpr int kGeneratorWarps = kThreads / 32 - 2;
constexpr int kGeneratorThreads = kGeneratorWarps * 32;
constexpr int kTmemColumns = 512;
constexpr uint32_t kOnes = 0x3f803f80u;                       // two bf16 ones

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
// kPair selects the two-SM forms (cta_group::2): one MMA spans the CTA pair of a cluster, M = 256.
template <bool kPair>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    if constexpr (kPair)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
}
// TMA load whose completion bytes are counted on the LEADER CTA's barrier (shared::cluster address).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1,
                                                 uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {   // possibly a peer's barrier
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquires peers' writes
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {   // arrives on both CTAs of the pair
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}
// One lane of a converged warp.  The TMA and MMA loops are run by their WHOLE warp with only the
// asynchronous instruction itself under this predicate: inside an `if (lane == 0)` region the
// compiler cannot prove the descriptors warp-uniform and wraps every UTCHMMA / UTMALDG in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY retry loop (~150 cycles per MMA, twice the MMA itself).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
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
__device__ __forceinline__ uint32_t cluster_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Programmatic dependent launch: let the next kernel of the stream (the split-K reduction) be
// scheduled while this grid drains / wait until the previous grid has completed and flushed.
__device__ __forceinline__ void launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void wait_for_primary() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
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
}
// The registers of an asynchronous TMEM read hold data only after the wait: passing them through
// the wait (and a second, empty statement ordered after it) keeps the compiler from using them early.
#define FEWBIT_SIXTEEN(v) "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), \
                          "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
__device__ __forceinline__ void tmem_load_wait(uint32_t (&v)[16], uint32_t (&w)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : FEWBIT_SIXTEEN(v)::"memory");
    asm volatile("" : FEWBIT_SIXTEEN(w)::"memory");
}
#undef FEWBIT_SIXTEEN

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100): start address, leading
// and stride byte offsets in 16-byte units, version 1, layout SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------------- random ----

// Philox4x32 with 7 rounds: the shortest variant that passes BigCrush (Salmon et al., "Parallel
// random numbers: as easy as 1, 2, 3", SC'11, table 2); cuRAND's 10 rounds add a safety margin a
// sketching matrix does not need, and the rounds are a third of the generators' instructions.
// -DFEWBIT_PHILOX_ROUNDS=10 restores it (S changes, nothing else does).
#ifndef FEWBIT_PHILOX_ROUNDS
#define FEWBIT_PHILOX_ROUNDS 7
#endif
constexpr int kPhiloxRounds = FEWBIT_PHILOX_ROUNDS;

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ uint4 operator()(uint4 c) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int round = 0; round < kPhiloxRounds; ++round) {
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
// N octets at once: the Philox chains and the Box-Muller evaluations are independent, and written
// as one straight-line block the scheduler interleaves them.
template <int N>
__device__ __forceinline__ void normal_octets(const Philox &rng, const uint32_t (&o)[N], const uint32_t (&p)[N],
                                              uint32_t off_lo, uint32_t off_hi, uint4 (&v)[N]) {
    uint4 c[N];
#pragma unroll
    for (int q = 0; q < N; ++q) c[q] = make_uint4(o[q], p[q], off_lo, off_hi);
    uint32_t a = rng.k0, b = rng.k1;
#pragma unroll
    for (int round = 0; round < kPhiloxRounds; ++round) {
#pragma unroll
        for (int q = 0; q < N; ++q) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c[q].x), lo0 = 0xD2511F53u * c[q].x;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[q].z), lo1 = 0xCD9E8D57u * c[q].z;
            c[q] = make_uint4(hi1 ^ c[q].y ^ a, lo1, hi0 ^ c[q].w ^ b, lo0);
        }
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
#pragma unroll
    for (int q = 0; q < N; ++q)
        v[q] = make_uint4(normal_pair(c[q].x), normal_pair(c[q].y), normal_pair(c[q].z), normal_pair(c[q].w));
}
// The same with the two halves taken from different slots: turns the counters `next` into Philox
// outputs (for the following slot) while `ready`, the outputs made one slot earlier, become normals.
template <int N>
__device__ __forceinline__ void philox_and_normals(const Philox &rng, uint4 (&next)[N], const uint4 (&ready)[N],
                                                   uint4 (&v)[N], uint32_t zero) {
    uint32_t a = rng.k0, b = rng.k1;
    uint32_t in[4 * N], out[4 * N];
#pragma unroll
    for (int q = 0; q < N; ++q) in[4 * q] = ready[q].x, in[4 * q + 1] = ready[q].y, in[4 * q + 2] = ready[q].z, in[4 * q + 3] = ready[q].w;
    // The Box-Muller pairs are dealt out between the Philox rounds.  ptxas would hoist all MUFU chains to
    // the top of the block (ALU idle, then XU idle); `zero` -- a kernel parameter that is 0 -- makes the
    // input of a pair nominally depend on the round before it, at one LOP3 per pair.
#pragma unroll
    for (int round = 0; round < kPhiloxRounds; ++round) {
#pragma unroll
        for (int q = 0; q < N; ++q) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, next[q].x), lo0 = 0xD2511F53u * next[q].x;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, next[q].z), lo1 = 0xCD9E8D57u * next[q].z;
            next[q] = make_uint4(hi1 ^ next[q].y ^ a, lo1, hi0 ^ next[q].w ^ b, lo0);
        }
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
#pragma unroll
        for (int i = 0; i < 4 * N; ++i)
            if (i >= 4 * N * round / kPhiloxRounds && i < 4 * N * (round + 1) / kPhiloxRounds)
                out[i] = normal_pair(in[i] ^ (next[i % N].x & zero));
    }
#pragma unroll
    for (int q = 0; q < N; ++q) v[q] = make_uint4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
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

__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct Params {
    float *out;          // [P, D] (split_k == 1; bf16 if out_bf16) or fp32 partials [split_k, P, D]
    int out_bf16;        // the result is written as bf16
    int ones_row;        // this row of S is all ones (its output row = column sums of X), -1: none
    int64_t tokens;      // N
    int features;        // D
    int rows;            // P
    int block_rows;      // BN
    int kblocks_per_split;
    int split_k;
    float scale;         // applied here only when split_k == 1
    uint32_t seed_lo, seed_hi, off_lo, off_hi;
    int kind;
    uint32_t zero;       // 0, opaque to the compiler (see philox_and_normals)
    int cluster_y;       // CTAs along grid.y that share one generated S slot
    int s_slots;         // S ring: entries of 128 tokens x (my share of) BN rows
    int s_tile_bytes;    // one 64-token half of an S slot
    unsigned long long *trace;   // FEWBIT_B200_SKETCH_TRACE: per-role time stamps of CTA (0,0,0), else null
    int debug;           // timing experiments only (results are garbage): 1 = skip generating S, 4 = skip issuing MMAs
};

template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
sketch_kernel(const __grid_constant__ CUtensorMap x_map, const Params prm) {
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled tiles need 1024-byte alignment.
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int slots = prm.s_slots;
    const uint32_t tile_bytes = (uint32_t)prm.s_tile_bytes;          // one 64-token half of an S slot
    uint8_t *s_ring = smem + kStages * kXStageBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_ring + slots * 2 * tile_bytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 2 * kMaxSlots + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool traced = prm.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    if (traced && threadIdx.x == 0) prm.trace[0] = now_ns();
    // Pair mode swaps the grid axes: a CTA pair must be adjacent along x of the cluster (2 x 1).
    const int p0 = (kPair ? blockIdx.y : blockIdx.x) * prm.block_rows;
    const int d0 = (kPair ? blockIdx.x : blockIdx.y) * kFeaturesPerCta;
    const int bn = prm.block_rows;
    const int nblocks = min(3, (prm.features - d0 + 127) / 128);   // 128-feature MMA blocks
    const int nboxes = min(6, (prm.features - d0 + 63) / 64);
    const int64_t total_kb = (prm.tokens + kBlockK - 1) / kBlockK;
    const int64_t kb_begin = (int64_t)blockIdx.z * prm.kblocks_per_split;      // even: S slots span two stages
    const int64_t kb_end = min(total_kb, kb_begin + prm.kblocks_per_split);
    const int iters = (int)max((int64_t)0, kb_end - kb_begin);

    auto full_x = [&](int s) { return smem_addr(bars + s); };
    auto empty_x = [&](int s) { return smem_addr(bars + kStages + s); };
    auto full_s = [&](int j) { return smem_addr(bars + 2 * kStages + j); };
    auto empty_s = [&](int j) { return smem_addr(bars + 2 * kStages + kMaxSlots + j); };
    const uint32_t accum_full = smem_addr(bars + 2 * kStages + 2 * kMaxSlots);
    const uint32_t x_ring = smem_addr(smem);

    const int cy = prm.cluster_y;                            // CTAs that share one generated S slot
    const uint32_t ry = kPair ? cluster_cta_x() : (cy > 1 ? cluster_cta_y() : 0);   // which share of S is mine
    const int my_rows = bn / cy;                             // rows of each S slot this CTA generates
    const bool pushing = !kPair && cy > 1;                   // my block also goes to the y-peer's shared memory
    // Pair mode (cluster 2 x 1, cta_group::2): the two CTAs own the two 384-feature halves of a
    // 768-feature slab and HALF of every S slot each; one M = 256 MMA, issued by the leader (ry = 0),
    // reads X^T from both shared memories and each half of S once for both.  The leader's full_x
    // counts the TMA bytes of both CTAs, its full_s the generator warps of both.
    const bool leader = !kPair || ry == 0;
    // the CTAs that share S are ranks [group0, group0 + Cy) of the cluster (group0 = 0 today; kept
    // general so that the cluster may grow along another axis)
    const uint32_t group0 = cluster_cta_rank() - ry;
    const uint16_t group_mask = (uint16_t)(((1u << (kPair ? 2 : cy)) - 1u) << group0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_x(s), 1);
            mbar_init(empty_x(s), kPair ? 1 : cy);                      // one commit per issuing CTA
        }
        for (int j = 0; j < slots; ++j) {
            // own generator warps, + the peer's forwarded arrival (pair leader) or the expect_tx arrival (pushing)
            mbar_init(full_s(j), kGeneratorWarps + ((kPair ? ry == 0 : cy > 1) ? 1 : 0));
            mbar_init(empty_s(j), kPair ? 1 : cy);
        }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM allocation is warp-wide; the same warp frees it (one warp per CTA, also in pair mode)
        if constexpr (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_addr(tmem_slot)),
                         "r"(kTmemColumns));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_addr(tmem_slot)),
                         "r"(kTmemColumns));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // Peers' barriers must exist before anyone signals them.  Generator warps that only touch their own
    // CTA (every mode but `pushing`) start at once and complete this cluster barrier after their loop.
    const bool clustered = kPair || cy > 1;
    const bool late_wait = clustered && !pushing && warp >= 2;
    if (clustered) {
        cluster_arrive();
        if (!late_wait) cluster_wait();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (traced && threadIdx.x == 0) prm.trace[1] = now_ns();

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer ----
        {
            for (int it = 0; it < iters; ++it) {
                const int s = it % kStages;
                mbar_wait(empty_x(s), ((it / kStages) & 1) ^ 1);
                const int token = (int)((kb_begin + it) * kBlockK);
                if (elect_one()) {
                    const uint32_t dst = x_ring + s * kXStageBytes;
                    if constexpr (kPair) {
                        if (leader) mbar_expect_tx(full_x(s), 2 * nboxes * kBoxBytes);     // both CTAs' boxes
                        const uint32_t bar = map_to_cta(full_x(s), group0);                 // counted on the leader's barrier
                        for (int b = 0; b < nboxes; ++b)
                            tma_load_2d_pair(dst + b * kBoxBytes, &x_map, d0 + 64 * b, token, bar);
                    } else {
                        mbar_expect_tx(full_x(s), nboxes * kBoxBytes);
                        for (int b = 0; b < nboxes; ++b)
                            tma_load_2d(dst + b * kBoxBytes, &x_map, d0 + 64 * b, token, full_x(s));
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer ----
        if (leader) {   // the whole warp walks the pipeline; one elected lane issues (see elect_one)
            // cute::UMMA::InstrDescriptor: D = f32, A = B = bf16, A MN-major, B K-major, N, M = 128
            // per CTA (256 across the pair).
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) |
                                   ((uint32_t)(bn >> 3) << 17) | (((kPair ? 256u : 128u) >> 4) << 24);
            // Shared-memory descriptors differ only in the 14-bit address field of their low word:
            //   A: 64-feature groups 8192 B apart (LBO), 8-token groups 1024 B apart (SBO); +2048 B per
            //      16 tokens; 128 features = two boxes.
            //   B: rows of 128 B (64 tokens), 8-row groups 1024 B apart (SBO); +32 B per 16 tokens.
            const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t a_lo = ((x_ring & 0x3FFFFu) >> 4) | ((uint32_t)(kBoxBytes >> 4) << 16);
            const uint32_t b_lo = ((smem_addr(s_ring) & 0x3FFFFu) >> 4) | ((16u >> 4) << 16);
            const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base, 0);
            const bool skip_mma = (prm.debug & 4) != 0;
            unsigned long long wait_x = 0, wait_s = 0;
            int j = 0;
            uint32_t sphase = 0;
            for (int it = 0; it < iters; ++it) {
                const int s = it % kStages, half = it & 1;
                const unsigned long long w0 = traced ? now_ns() : 0;
                if (half == 0) {
                    if constexpr (kPair) mbar_wait_cluster(full_s(j), sphase);   // the peer's generators wrote its half
                    else mbar_wait(full_s(j), sphase);
                }
                const unsigned long long w1 = traced ? now_ns() : 0;
                mbar_wait(full_x(s), (it / kStages) & 1);
                if (traced && lane == 0) {
                    const unsigned long long w2 = now_ns();
                    wait_s += w1 - w0, wait_x += w2 - w1;
                    if (it == 0) prm.trace[2] = w2;
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool slot_done = half == 1 || it == iters - 1;      // both halves of the S slot consumed
                if (elect_one()) {
                    if (!skip_mma) {
                        const uint32_t a_stage = a_lo + (uint32_t)s * (uint32_t)(kXStageBytes >> 4);
                        const uint32_t b_tile = b_lo + (uint32_t)(((uint32_t)(2 * j + half) * tile_bytes) >> 4);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t desc_b = ((uint64_t)desc_hi << 32) | (b_tile + 2u * k);
#pragma unroll
                            for (int m = 0; m < 3; ++m) {
                                if (m < nblocks) {
                                    const uint64_t desc_a = ((uint64_t)desc_hi << 32) |
                                                            (a_stage + (uint32_t)((m * 2 * kBoxBytes + 2048 * k) >> 4));
                                    umma_bf16<kPair>(tmem + m * kMaxRows, desc_a, desc_b, idesc, (it | k) != 0);
                                }
                            }
                        }
                    }
                    // The X stage is reusable once these MMAs have read it -- in every CTA of the cluster;
                    // the S slot once both of its halves have been.
                    if constexpr (kPair) {
                        umma_commit_pair(empty_x(s), group_mask);
                        if (slot_done) umma_commit_pair(empty_s(j), group_mask);
                    } else if (cy > 1) {
                        umma_commit_cluster(empty_x(s), group_mask);
                        if (slot_done) umma_commit_cluster(empty_s(j), group_mask);
                    } else {
                        umma_commit(empty_x(s));
                        if (slot_done) umma_commit(empty_s(j));
                    }
                }
                __syncwarp();
                if (slot_done && ++j == slots) j = 0, sphase ^= 1;
            }
            if (elect_one()) {   // accumulators complete (in both CTAs of a pair)
                if constexpr (kPair) umma_commit_pair(accum_full, group_mask);
                else umma_commit(accum_full);
            }
            __syncwarp();
            if (traced && lane == 0) prm.trace[3] = now_ns(), prm.trace[6] = wait_x, prm.trace[7] = wait_s;
        } else {
            // Pair mode, second CTA: forward "my half of the S slot is written" to the leader's barrier.
            // The cluster-scope release costs ~0.5 us; paid here, on an otherwise idle warp, it stays
            // off the generators' critical path.
            int j = 0;
            uint32_t sphase = 0;
            for (int n = 0; n < (iters + 1) / 2; ++n) {
                mbar_wait(full_s(j), sphase);
                if (elect_one()) mbar_arrive_cluster(map_to_cta(full_s(j), group0));
                __syncwarp();
                if (++j == slots) j = 0, sphase ^= 1;
            }
        }
    } else {
        // -------------------------------------------------------------- generators ----
        // One S slot = 128 tokens = two K-major SW128 tiles of 64 tokens; the generators run up to
        // `slots` slots ahead of the MMAs, independently of the X ring.
        const Philox rng{prm.seed_lo, prm.seed_hi};
        const int gt = threadIdx.x - 64;
        unsigned long long wait_e = 0, busy = 0;
        long long ph[4] = {0, 0, 0, 0};      // cycles: generate + store, proxy fence, block sync + push, arrive
        const int row0 = (int)ry * my_rows;                // this CTA's block of the slot
        const int place = kPair ? 0 : row0;                // pair mode: the block sits at the tile start
        // Gaussian: one Philox call = 8 normals = one 16-byte chunk, 16 chunks per row and slot.  The
        // chunks are dealt evenly: `per` to each of the first `active` threads, taken in groups of up to
        // four whose Philox chains and Box-Muller evaluations interleave (a single chain is latency-
        // bound: ~10 dependent multiply-xor rounds, then lg2 -> sqrt).
        const int chunks = my_rows * 16;
        const int per = (chunks + kGeneratorThreads - 1) / kGeneratorThreads;
        const int active = (chunks + per - 1) / per;
        const int groups = (per + 3) / 4, group = (per + groups - 1) / groups;
        const int nslots = (iters + 1) / 2;
        // One pass over the S ring; fill(slot, kb) writes this CTA's block of the slot starting at k-block kb.
        auto for_each_slot = [&](auto &&fill) {
            int j = 0;
            uint32_t sphase = 0;
            for (int n = 0; n < nslots; ++n) {
                const unsigned long long w0 = traced && gt == 0 ? now_ns() : 0;
                mbar_wait(empty_s(j), sphase ^ 1);
                const unsigned long long g0 = traced && gt == 0 ? now_ns() : 0;
                const long long c0 = traced && gt == 0 ? clock64() : 0;
                if (traced && gt == 0) wait_e += g0 - w0;
                if (pushing && gt == 0) mbar_expect_tx(full_s(j), (uint32_t)((cy - 1) * my_rows * 128 * 2));   // the y-peer's block
                uint8_t *slot = s_ring + (size_t)j * 2 * tile_bytes;
                if (!(prm.debug & 1)) fill(slot, kb_begin + 2 * n);
                const long long c1 = traced && gt == 0 ? clock64() : 0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic -> async proxy
                const long long c2 = traced && gt == 0 ? clock64() : 0;
                if (pushing) {
                    // all generator threads have written (and fenced) this CTA's block: push both halves to the y-peers
                    asm volatile("bar.sync 1, %0;" ::"n"(kGeneratorThreads) : "memory");
                    if (gt == 0) {
                        for (uint32_t y = 0; y < (uint32_t)cy; ++y)
                            if (y != ry)
                                for (uint32_t h = 0; h < 2; ++h) {
                                    const uint32_t block = smem_addr(slot) + h * tile_bytes + row0 * 128;
                                    bulk_copy_to_peer(map_to_cta(block, group0 + y), block, my_rows * 128, map_to_cta(full_s(j), group0 + y));
                                }
                    }
                }
                __syncwarp();
                const long long c3 = traced && gt == 0 ? clock64() : 0;
                if (traced && gt == 0) ph[0] += c1 - c0, ph[1] += c2 - c1, ph[2] += c3 - c2;
                if (traced && gt == 0) busy += now_ns() - g0;
                if (lane == 0) mbar_arrive(full_s(j));       // pair mode, second CTA: warp 1 forwards it
                if (traced && gt == 0) ph[3] += clock64() - c3;
                if (++j == slots) j = 0, sphase ^= 1;
            }
        };
        // row r of the tile, 16-byte chunk o of the slot's 256-byte row (K-major SW128: chunk XOR row mod 8)
        auto chunk_at = [&](uint8_t *slot, int r, int o) {
            return slot + (o >> 3) * tile_bytes + (r >> 3) * 1024 + (r & 7) * 128 + (((o & 7) ^ (r & 7)) << 4);
        };
        // Gaussian, up to four chunks per thread: a thread's chunks sit at the same (row, chunk) of every
        // slot, and the Philox calls of the NEXT slot are made in the same straight-line block as the
        // Box-Muller evaluations of this one -- integer multiply-xor rounds and MUFU chains interleave,
        // where one after the other they leave the ALU and the XU pipes idle in turn.
        auto gaussian_pipelined = [&](auto count) {
            constexpr int kPer = decltype(count)::value;
            const bool working = gt < active;
            uint32_t o[kPer], p[kPer];
            uint4 c[kPer];
            int at[kPer];
#pragma unroll
            for (int q = 0; q < kPer; ++q) {
                at[q] = min(gt + q * active, chunks - 1);        // a duplicate at the ragged end
                o[q] = (uint32_t)(at[q] & 15), p[q] = (uint32_t)(p0 + row0 + (at[q] >> 4));
                c[q] = rng(make_uint4((uint32_t)(kb_begin * 8) + o[q], p[q], prm.off_lo, prm.off_hi));
            }
            for_each_slot([&](uint8_t *slot, int64_t kb) {
                if (!working) return;
                uint4 v[kPer], next[kPer];
#pragma unroll
                for (int q = 0; q < kPer; ++q) next[q] = make_uint4((uint32_t)((kb + 2) * 8) + o[q], p[q], prm.off_lo, prm.off_hi);
                philox_and_normals<kPer>(rng, next, c, v, prm.zero);
#pragma unroll
                for (int q = 0; q < kPer; ++q) {
                    if ((int)p[q] == prm.ones_row) v[q] = make_uint4(kOnes, kOnes, kOnes, kOnes);
                    *reinterpret_cast<uint4 *>(chunk_at(slot, place + (at[q] >> 4), at[q] & 15)) = v[q];
                    c[q] = next[q];
                }
            });
        };
        if (prm.kind == 0 && per <= 4) {
            switch (per) {
                case 1: gaussian_pipelined(std::integral_constant<int, 1>{}); break;
                case 2: gaussian_pipelined(std::integral_constant<int, 2>{}); break;
                case 3: gaussian_pipelined(std::integral_constant<int, 3>{}); break;
                default: gaussian_pipelined(std::integral_constant<int, 4>{}); break;
            }
        } else if (prm.kind == 0) {
            // more than four chunks per thread (BN rows unshared): groups of up to four interleaved chains
            for_each_slot([&](uint8_t *slot, int64_t kb) {
                if (gt >= active) return;
                auto run = [&](auto count, int q0) {
                    constexpr int kCount = decltype(count)::value;
                    uint32_t o[kCount], p[kCount];
                    uint4 v[kCount];
                    int at[kCount];
#pragma unroll
                    for (int q = 0; q < kCount; ++q) {
                        at[q] = min(gt + (q0 + q) * active, chunks - 1);
                        o[q] = (uint32_t)(kb * 8 + (at[q] & 15)), p[q] = (uint32_t)(p0 + row0 + (at[q] >> 4));
                    }
                    normal_octets<kCount>(rng, o, p, prm.off_lo, prm.off_hi, v);
#pragma unroll
                    for (int q = 0; q < kCount; ++q) {
                        if ((int)p[q] == prm.ones_row) v[q] = make_uint4(kOnes, kOnes, kOnes, kOnes);
                        *reinterpret_cast<uint4 *>(chunk_at(slot, place + (at[q] >> 4), at[q] & 15)) = v[q];
                    }
                };
                for (int q0 = 0; q0 < per; q0 += group) {
                    switch (min(group, per - q0)) {
                        case 1: run(std::integral_constant<int, 1>{}, q0); break;
                        case 2: run(std::integral_constant<int, 2>{}, q0); break;
                        case 3: run(std::integral_constant<int, 3>{}, q0); break;
                        default: run(std::integral_constant<int, 4>{}, q0); break;
                    }
                }
            });
        } else {
            // One Philox call = 128 signs = one row of the slot.  A task expands one of its four words
            // (32 tokens = four chunks); the four tasks of a row repeat the call in neighbouring lanes,
            // which costs nothing in a SIMT warp and keeps 4 x rows threads busy.
            for_each_slot([&](uint8_t *slot, int64_t kb) {
                for (int task = gt; task < 4 * my_rows; task += kGeneratorThreads) {
                    const int row = task >> 2, q = task & 3;
                    const uint4 w = sign_block(rng, (uint32_t)(kb >> 1), (uint32_t)(p0 + row0 + row), prm.off_lo, prm.off_hi);
                    const uint32_t word = q == 0 ? w.x : q == 1 ? w.y : q == 2 ? w.z : w.w;
                    const bool ones = p0 + row0 + row == prm.ones_row;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t bits = word >> (c * 8);
                        *reinterpret_cast<uint4 *>(chunk_at(slot, place + row, q * 4 + c)) =
                            ones ? make_uint4(kOnes, kOnes, kOnes, kOnes)
                                 : make_uint4(sign_pair(bits), sign_pair(bits >> 2), sign_pair(bits >> 4), sign_pair(bits >> 6));
                    }
                }
            });
        }
        if (late_wait) cluster_wait();
        if (traced && gt == 0) prm.trace[8] = wait_e, prm.trace[9] = now_ns(), prm.trace[10] = busy;
        if (traced && gt == 0)
            for (int q = 0; q < 4; ++q) prm.trace[11 + q] = (unsigned long long)(ph[q] / max(nslots, 1));
    }
    // -------------------------------------------------------------------- epilogue ----
    // All warps: a warp reads the TMEM lanes of its quadrant (warp % 4); the warps of a
    // quadrant share its (feature block, 16-column) units.  For a fixed sketch row the 32 lanes
    // hold 32 consecutive features: every store instruction writes one 128-byte line.
    mbar_wait(accum_full, 0);
    if (prm.split_k > 1) launch_dependents();
    if (traced && warp == 4 && lane == 0) prm.trace[4] = now_ns();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quarter = warp & 3;                       // TMEM lanes [32q, 32q + 32)
    const int units_per_block = bn / 16;
    auto load_unit = [&](int m, int c, uint32_t (&v)[16]) {      // 16 sketch rows of this lane's feature (asynchronous)
        tmem_load16(tmem_base + ((uint32_t)(quarter * 32) << 16) + m * kMaxRows + c * 16, v);
    };
    const bool narrow = prm.out_bf16 && prm.split_k == 1;        // partials are always fp32
    auto store_unit = [&](float *out, int m, int c, const uint32_t (&v)[16], float scale) {
        const int d = d0 + m * 128 + quarter * 32 + lane;
        if (d < prm.features) {
            const int64_t at = (int64_t)(p0 + c * 16) * prm.features + d;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (p0 + c * 16 + q >= prm.rows) continue;
                const float value = iters ? __uint_as_float(v[q]) * scale : 0.0f;
                if (narrow) reinterpret_cast<__nv_bfloat16 *>(out)[at + (int64_t)q * prm.features] = __float2bfloat16_rn(value);
                else out[at + (int64_t)q * prm.features] = value;
            }
        }
    };
    {
        // split_k == 1: the result; else this split's partial, added up by reduce_splits_kernel
        float *out = prm.out + (prm.split_k > 1 ? (int64_t)blockIdx.z * prm.rows * prm.features : 0);
        const float scale = prm.split_k > 1 ? 1.0f : prm.scale;
        // two units in flight per warp: the second TMEM read overlaps the first unit's stores
        const int total = nblocks * units_per_block, step = kThreads / 128;
        for (int t = warp >> 2; t < total; t += 2 * step) {
            uint32_t v[16], w[16] = {};
            const bool two = t + step < total;
            load_unit(t / units_per_block, t % units_per_block, v);
            if (two) load_unit((t + step) / units_per_block, (t + step) % units_per_block, w);
            tmem_load_wait(v, w);
            store_unit(out, t / units_per_block, t % units_per_block, v, scale);
            if (two) store_unit(out, (t + step) / units_per_block, (t + step) % units_per_block, w, scale);
        }
    }
    if (traced && warp == 4 && lane == 0) prm.trace[5] = now_ns();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (clustered) cluster_sync();            // nobody leaves while peers may still signal it
    if (warp == 1) {
        if constexpr (kPair)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"(kTmemColumns));
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"(kTmemColumns));
    }
}

// out = scale * sum of the split-K partials (fixed order: deterministic), fp32 or bf16.  kVec = 4 when the
// buffers allow 16-byte loads.
template <int kVec, typename Out>
__global__ void reduce_splits_kernel(const float *partials, Out *out, int64_t count, int splits, float scale) {
    using Vec = typename std::conditional<kVec == 4, float4, float>::type;
    const int64_t n = count / kVec;
    wait_for_primary();              // launched early (programmatic stream serialisation): the partials are complete from here
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        Vec acc = reinterpret_cast<const Vec *>(partials)[i];
        for (int s = 1; s < splits; ++s) {
            const Vec v = __ldcs(reinterpret_cast<const Vec *>(partials + (int64_t)s * count) + i);
            if constexpr (kVec == 4) acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
            else acc += v;
        }
        if constexpr (kVec == 4) {
            acc.x *= scale, acc.y *= scale, acc.z *= scale, acc.w *= scale;
            if constexpr (std::is_same<Out, float>::value) {
                reinterpret_cast<float4 *>(out)[i] = acc;
            } else {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
                reinterpret_cast<uint2 *>(out)[i] = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
            }
        } else {
            out[i] = static_cast<Out>(acc * scale);
        }
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
    // a process-wide driver entry point; C++11 makes the one-time initialisation thread-safe
    static const EncodeTiled fn = [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &status) == cudaSuccess &&
            status == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiled>(ptr);
        return static_cast<EncodeTiled>(nullptr);
    }();
    return fn;
}

// Pick how CTAs share a generated S slot (Cy feature tiles, or a cta_group::2 pair), BN (multiple of
// 16 and of 8 Cy, <= 160) and split_k: minimise waves * (time of one CTA).  Measured on B200
// (profiles/r01_sketch_kernel.md): clusters of 8 along y lose to independent CTAs at D = 3072 (8-CTA
// placement leaves SMs idle and the lock-step `empty` barriers couple eight pipelines), and sharing X
// between row tiles by TMA multicast loses too (L2 is not the limiter), so Cy <= 2 and X is not shared.
struct Plan {
    int bn, split_k, cy;
    bool pair;
    int kblocks_per_split;       // even: an S slot spans two 64-token stages
    int s_slots, s_tile_bytes, smem_bytes;
};

static int env_int(const char *name, int fallback) {
    const char *env = std::getenv(name);
    return env ? std::atoi(env) : fallback;
}

static Plan plan(int rows, int features, int64_t tokens, int kind, int sms) {
    Plan pl{};
    const int dtiles = (features + kFeaturesPerCta - 1) / kFeaturesPerCta;
    const int64_t kblocks = std::max<int64_t>(1, (tokens + kBlockK - 1) / kBlockK);
    // Default from the A/B runs in profiles/r01_sketch_kernel.md: S sharing between the two feature
    // tiles of D = 768 helps (133 -> 123 us).
    pl.cy = dtiles == 2 ? 2 : 1;
    // Pair mode (cta_group::2 MMAs over a 2 x 1 cluster) whenever the feature tiles come in full
    // pairs: each S slot is generated once per pair and read from shared memory once per pair.
    // Measured (profiles/r01_sketch_kernel.md): with Gaussian entries the kernel is bound by generating
    // S and the pair halves that work (D = 3072: 493 -> 406 us); with Rademacher entries it is bound by
    // the MMA pipeline, and the extra hop of the peer's "S ready" signal costs more than it saves.
    pl.pair = features % (2 * kFeaturesPerCta) == 0 && env_int("FEWBIT_B200_SKETCH_PAIR", 1) != 0;   // 0: A/B runs
    if (pl.pair) pl.cy = 2;
    if (const int v = env_int("FEWBIT_B200_SKETCH_CLUSTER", 0)) {   // A/B runs: Cy without cta_group::2
        if (v >= 1 && v <= 2) pl.cy = std::min(pl.cy, v);
        pl.pair = false;
    }
    double best = 1e300;
    pl.bn = 64, pl.split_k = 1;
    const int only_bn = env_int("FEWBIT_B200_SKETCH_BN", 0), only_sk = env_int("FEWBIT_B200_SKETCH_SPLITK", 0);   // tuning runs
    for (int cand = 160; cand >= 64; cand -= 16) {
        if ((cand / 8) % pl.cy != 0) continue;
        if (only_bn && cand != only_bn) continue;
        const int ptiles = (rows + cand - 1) / cand;
        for (int sk = 1; sk <= 8 && sk <= kblocks; ++sk) {
            if (only_sk && sk != only_sk) continue;
            const int64_t ctas = (int64_t)ptiles * dtiles * sk;
            const int64_t waves = (ctas + sms - 1) / sms;
            // Cycles, fitted to the (BN, split_k) sweep in profiles/r01_sketch_kernel.md.  Per 64-token
            // stage: ~17 cycles per generated Gaussian row (the generators are latency-bound), never
            // under ~1300 (pipeline round trip; 12 MMAs of ~90-107 cycles).  Per CTA: ~20000 for
            // prologue + epilogue (the epilogue writes at HBM speed).  Split-K adds the partials'
            // round trip through memory: 8 bytes per output element and split at ~5 TB/s.
            const double generated = kind == 0 ? (double)cand / pl.cy * 17.0 : 0.0;
            const double block = std::max({generated, 1300.0, cand * 9.0});   // 12 MMAs: ~9 cycles per row
            const double per_cta = (double)((kblocks + sk - 1) / sk) * block + 20000.0;
            const double reduce = sk > 1 ? (double)sk * rows * features * 8.0 / 5e12 * 1.9e9 : 0.0;
            const double cost = (double)waves * per_cta + reduce;
            if (cost < best) best = cost, pl.bn = cand, pl.split_k = sk;
        }
    }
    int per = (int)((kblocks + pl.split_k - 1) / pl.split_k);
    pl.kblocks_per_split = per + (per & 1);
    // Shared memory: three X stages, the rest holds S slots (each CTA stores only the rows it generates
    // in pair mode, the whole BN-row slot otherwise).
    const int tile_rows = ((pl.pair ? pl.bn / 2 : pl.bn) + 7) / 8 * 8;
    pl.s_tile_bytes = tile_rows * 128;
    const int room = kSmemLimit - 1024 /* alignment */ - kBarrierBytes - kStages * kXStageBytes;
    pl.s_slots = std::min(kMaxSlots, room / (2 * pl.s_tile_bytes));
    if (const int v = env_int("FEWBIT_B200_SKETCH_SLOTS", 0)) pl.s_slots = std::min(pl.s_slots, std::max(v, 1));
    pl.smem_bytes = kStages * kXStageBytes + pl.s_slots * 2 * pl.s_tile_bytes + kBarrierBytes + 1024;
    return pl;
}

}  // namespace sketch
}  // namespace fewbit

using namespace fewbit;
using namespace fewbit::sketch;

extern "C" {

size_t fewbit_sketch_workspace_bytes(int64_t tokens, int features, int rows) {
    size_t bytes = 0;
    for (int kind = 0; kind < 2; ++kind) {      // the larger of the two kinds
        const Plan pl = plan(rows, features, tokens, kind, sm_count());
        if (pl.split_k > 1) bytes = std::max(bytes, (size_t)pl.split_k * rows * features * sizeof(float));
    }
    return bytes;
}

int fewbit_sketch_forward(const void *x, float *out, void *workspace, int64_t tokens, int features,
                          int rows, int kind, float scale, uint64_t seed, uint64_t offset, void *stream) {
    return fewbit_sketch_project(x, out, FEWBIT_F32, workspace, tokens, features, rows, 0, kind, scale, seed, offset, stream);
}

int fewbit_sketch_project(const void *x, void *out_any, int out_dtype, void *workspace, int64_t tokens, int features,
                          int sketch_rows, int column_sums, int kind, float scale, uint64_t seed, uint64_t offset,
                          void *stream) {
    if (tokens < 0 || features <= 0 || sketch_rows <= 0 || (kind != 0 && kind != 1)) return FEWBIT_EINVAL;
    if (out_dtype != FEWBIT_F32 && out_dtype != FEWBIT_BF16) return FEWBIT_EDTYPE;
    float *out = static_cast<float *>(out_any);
    const int rows = sketch_rows + (column_sums ? 1 : 0);      // the extra row of S is all ones
    if (!x || !out) return FEWBIT_EINVAL;
    if (features % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15)) return FEWBIT_EALIGN;  // TMA strides
    EncodeTiled encode = encode_tiled();
    if (!encode) return (int)cudaErrorNotSupported;
    cudaStream_t s = (cudaStream_t)stream;
    const Plan pl = plan(rows, features, tokens, kind, sm_count());
    const int bn = pl.bn, split_k = pl.split_k, cy = pl.cy;
    const bool pair = pl.pair;
    if (pl.s_slots < 1) return (int)cudaErrorInvalidConfiguration;

    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)features, (cuuint64_t)std::max<int64_t>(tokens, 1)};
    const cuuint64_t strides[1] = {(cuuint64_t)features * 2};
    const cuuint32_t box[2] = {64, 64}, elem[2] = {1, 1};
    if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(x), dims, strides, box, elem,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return (int)cudaErrorInvalidValue;

    Params prm;
    prm.out = split_k > 1 ? static_cast<float *>(workspace) : out;
    prm.tokens = tokens, prm.features = features, prm.rows = rows, prm.block_rows = bn;
    prm.kblocks_per_split = pl.kblocks_per_split;
    prm.split_k = split_k, prm.scale = scale, prm.kind = kind, prm.cluster_y = cy;
    prm.s_slots = pl.s_slots, prm.s_tile_bytes = pl.s_tile_bytes, prm.zero = 0;
    prm.out_bf16 = out_dtype == FEWBIT_BF16, prm.ones_row = column_sums ? sketch_rows : -1;
    prm.debug = env_int("FEWBIT_B200_SKETCH_DEBUG", 0);
    prm.trace = nullptr;
    static unsigned long long *trace_buffer = nullptr;
    const bool tracing = std::getenv("FEWBIT_B200_SKETCH_TRACE") != nullptr;
    if (tracing) {
        if (!trace_buffer) cudaMalloc(&trace_buffer, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(trace_buffer, 0, 16 * sizeof(unsigned long long), s);
        prm.trace = trace_buffer;
    }
    prm.seed_lo = (uint32_t)seed, prm.seed_hi = (uint32_t)(seed >> 32);
    prm.off_lo = (uint32_t)offset, prm.off_hi = (uint32_t)(offset >> 32);

    // function attributes live in the device's context: set once per device
    static bool configured[64] = {};
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64) device = 0;
    if (!configured[device]) {
        cudaError_t e = cudaFuncSetAttribute(sketch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(sketch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) return (int)e;
        configured[device] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((rows + bn - 1) / bn, (features + kFeaturesPerCta - 1) / kFeaturesPerCta, split_k);
    if (pair) std::swap(cfg.gridDim.x, cfg.gridDim.y);      // feature tiles (the pairs) along x
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair ? 2 : 1, attr[0].val.clusterDim.y = pair ? 1 : cy, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    if (split_k > 1 && !workspace) return FEWBIT_EINVAL;
    cudaError_t launched = pair ? cudaLaunchKernelEx(&cfg, sketch_kernel<true>, map, prm)
                                : cudaLaunchKernelEx(&cfg, sketch_kernel<false>, map, prm);
    if (launched != cudaSuccess) return (int)launched;
    note_launch();
    if (tracing) {   // diagnostics only: synchronises
        unsigned long long t[16];
        cudaStreamSynchronize(s);
        cudaMemcpy(t, trace_buffer, sizeof(t), cudaMemcpyDeviceToHost);
        std::fprintf(stderr,
                     "[sketch trace] D=%d bn=%d split_k=%d pair=%d slots=%d grid=%ux%ux%u | setup %.1f us, first MMA +%.1f, "
                     "MMA loop %.1f (waited X %.1f, S %.1f), generators done +%.1f (waited empty %.1f, generating %.1f), "
                     "accumulators seen +%.1f, epilogue %.1f | generator thread 0, cycles per 128-token slot: generate %llu, proxy fence %llu, "
                     "sync+push %llu, arrive %llu\n",
                     features, bn, split_k, (int)pair, pl.s_slots, cfg.gridDim.x, cfg.gridDim.y, cfg.gridDim.z,
                     (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3, (t[3] - t[2]) * 1e-3, t[6] * 1e-3, t[7] * 1e-3,
                     (t[9] - t[1]) * 1e-3, t[8] * 1e-3, t[10] * 1e-3, (t[4] - t[1]) * 1e-3, (t[5] - t[4]) * 1e-3, t[11], t[12], t[13], t[14]);
    }
    if (split_k > 1) {
        const int64_t count = (int64_t)rows * features;
        const bool wide = count % 4 == 0 && ((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
        const unsigned blocks = (unsigned)std::min<int64_t>((count / (wide ? 4 : 1) + 255) / 256, sm_count() * 8);
        cudaLaunchConfig_t rcfg{};
        rcfg.gridDim = dim3(blocks), rcfg.blockDim = dim3(256), rcfg.stream = s;
        cudaLaunchAttribute early[1];
        early[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        early[0].val.programmaticStreamSerializationAllowed = 1;
        rcfg.attrs = early, rcfg.numAttrs = 1;
        const float *partials = static_cast<const float *>(workspace);
        __nv_bfloat16 *narrow = static_cast<__nv_bfloat16 *>(out_any);
        const cudaError_t reduced =
            prm.out_bf16 ? (wide ? cudaLaunchKernelEx(&rcfg, reduce_splits_kernel<4, __nv_bfloat16>, partials, narrow, count, split_k, scale)
                                 : cudaLaunchKernelEx(&rcfg, reduce_splits_kernel<1, __nv_bfloat16>, partials, narrow, count, split_k, scale))
                         : (wide ? cudaLaunchKernelEx(&rcfg, reduce_splits_kernel<4, float>, partials, out, count, split_k, scale)
                                 : cudaLaunchKernelEx(&rcfg, reduce_splits_kernel<1, float>, partials, out, count, split_k, scale));
        if (reduced != cudaSuccess) return (int)reduced;
        note_launch();
    }
    return (int)cudaGetLastError();
}

int fewbit_sketch_plan(int64_t tokens, int features, int rows, int kind, int sms, int out[8]) {
    if (tokens < 0 || features <= 0 || rows <= 0 || (kind != 0 && kind != 1) || sms <= 0 || !out) return FEWBIT_EINVAL;
    const Plan pl = plan(rows, features, tokens, kind, sms);
    out[0] = pl.bn, out[1] = pl.split_k, out[2] = pl.cy, out[3] = pl.pair, out[4] = pl.kblocks_per_split;
    out[5] = pl.s_slots, out[6] = pl.s_tile_bytes, out[7] = pl.smem_bytes;
    return FEWBIT_OK;
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
