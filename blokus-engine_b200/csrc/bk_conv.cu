// bk_conv.cu — hand-written sm_100a tensor-core kernel for the one dense contraction on the path: the
// 3x3 convolutions of the reference's policy/value ResNet trunk (model/resnet.py:8-26,51-52;
// SURVEY.md §8f row f2).  tcgen05.mma with the accumulator in TMEM, operands staged by TMA into
// 128B-swizzled shared memory, BatchNorm folded into weights/bias, bias + residual + ReLU fused into the
// epilogue that reads TMEM with tcgen05.ld.
//
// Formulation.  Activations live in HBM as a zero-padded NHWC matrix X[m][c], bf16, with
// m = image*441 + row*21 + col (row, col in 0..20; row 20 and col 20 are zero padding shared by neighbours),
// so a 3x3 tap (dy, dx) is a pure row shift of dy*21 + dx and the convolution is nine shifted GEMMs
//     Y[m][n] = sum_tap sum_k X[m + dy*21 + dx][k] * W[tap][n][k]
// whose A tiles are plain 2-D TMA boxes (out-of-range rows are zero-filled by TMA).  One CTA computes a
// 128-row x 256-channel output tile: 9 taps x 4 K-chunks of 64 = 36 pipeline steps of {A 16 KB, B 32 KB},
// each 4 UMMA instructions M128 N256 K16 into 256 TMEM columns.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
// lane), warps 2..5 = epilogue (each owns the 32 TMEM lanes of its warp-in-warpgroup rank).
#include "bk_host.h"

#ifndef BK_WARP_EMU
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>
#include <stdio.h>

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;
constexpr int kBlockK = 64;
constexpr int kChannels = 256;
constexpr int kStages = 4;
constexpr int kTaps = 9;
constexpr int kStepsPerTap = kChannels / kBlockK;          // 4
constexpr int kSteps = kTaps * kStepsPerTap;               // 36
constexpr uint32_t kBytesA = kBlockM * kBlockK * 2;        // 16 KB
constexpr uint32_t kBytesB = kBlockN * kBlockK * 2;        // 32 KB
constexpr uint32_t kBytesStage = kBytesA + kBytesB;
constexpr uint32_t kSmemBytes = kStages * kBytesStage + 256 + 1024;   // + barriers + alignment slack
constexpr uint32_t kTmemCols = 256;
constexpr int kPadDim = 21;
constexpr int kPadImage = kPadDim * kPadDim;               // 441

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major, 128B-swizzled operand tile whose rows are 128 bytes: 8-row groups are 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFFu);          // start address
    d |= uint64_t(1) << 16;                              // leading byte offset (unused for swizzled K-major)
    d |= uint64_t((1024u >> 4) & 0x3FFFu) << 32;         // stride byte offset
    d |= uint64_t(1) << 46;                              // descriptor version (sm_100)
    d |= uint64_t(2) << 61;                              // SWIZZLE_128B
    return d;
}
// D(f32) += A(bf16, K-major) * B(bf16, K-major)^T, M = 128, N = 256
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(kBlockN >> 3) << 17) | (uint32_t(kBlockM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kInstrDesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// Y = act(conv3x3(X) + bias [+ R]) on the padded NHWC layout; pad rows of Y are written as zeros.
// Persistent: cluster c computes tile groups c, c + n_clusters, ...; CTA r of the cluster takes tile group*CS + r.
// The accumulator is double-buffered in TMEM (2 x 256 columns), so the epilogue of tile i overlaps the TMA/MMA
// main loop of tile i + 1.  The weight tile of a step is the same for every CTA: each CTA of a cluster fetches
// 1/CS of it and TMA-multicasts that slice into all CS shared memories (L2 -> SM traffic per step drops from
// 48 KB to 16 + 32/CS KB); a stage is recycled when the MMAs of ALL CTAs of the cluster have consumed it.
template <int CS>
__global__ void __launch_bounds__(192, 1)
k_conv3x3_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
             const float* __restrict__ bias, const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
             int m_total, int n_tiles, int steps_per_tap, int relu) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + kStages * kBytesStage;
    auto full_bar = [&](int s) { return bar_base + 8u * uint32_t(s); };
    auto empty_bar = [&](int s) { return bar_base + 8u * uint32_t(kStages + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * uint32_t(2 * kStages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * uint32_t(2 * kStages + 2 + a); };
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + kStages * kBytesStage + 8 * (2 * kStages + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int steps = kTaps * steps_per_tap;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CS); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(2 * kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CS > 1) cluster_sync_all();                  // peers' barriers are initialised before anyone multicasts
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_holder;
    const int crank = CS > 1 ? int(cluster_ctarank()) : 0;
    const int n_groups = (n_tiles + CS - 1) / CS;
    const int first_group = int(blockIdx.x) / CS, group_stride = int(gridDim.x) / CS;
    constexpr uint16_t kMcMask = uint16_t((1u << CS) - 1u);
    constexpr uint32_t kSliceRows = kBlockN / CS, kSliceBytes = kBytesB / CS;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;                                           // pipeline step counter across tiles
            for (int grp = first_group; grp < n_groups; grp += group_stride) {
                const int m0 = (grp * CS + crank) * kBlockM;
                for (int ks = 0; ks < steps; ++ks, ++it) {
                    const int s = int(it % kStages);
                    const uint32_t ph = (it / kStages) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(full_bar(s), kBytesStage);
                    const int tap = ks / steps_per_tap, kc = ks - tap * steps_per_tap;
                    const int shift = (tap / 3 - 1) * kPadDim + (tap % 3 - 1);
                    const uint32_t a_dst = smem_base + uint32_t(s) * kBytesStage;
                    tma_load_2d(a_dst, &map_x, full_bar(s), kc * kBlockK, m0 + shift);
                    if (CS == 1) tma_load_2d(a_dst + kBytesA, &map_w, full_bar(s), kc * kBlockK, tap * kBlockN);
                    else tma_load_2d_mc(a_dst + kBytesA + uint32_t(crank) * kSliceBytes, &map_w, full_bar(s), kc * kBlockK,
                                        tap * kBlockN + crank * int(kSliceRows), kMcMask);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, n_acc = 0;
            for (int grp = first_group; grp < n_groups; grp += group_stride, ++n_acc) {
                const int acc = int(n_acc & 1u);
                mbar_wait(tmem_empty_bar(acc), ((n_acc >> 1) & 1u) ^ 1u);      // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + uint32_t(acc) * kTmemCols;
                for (int ks = 0; ks < steps; ++ks, ++it) {
                    const int s = int(it % kStages);
                    const uint32_t ph = (it / kStages) & 1u;
                    mbar_wait(full_bar(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_base + uint32_t(s) * kBytesStage;
                    const uint64_t da = umma_desc(a_addr), db = umma_desc(a_addr + kBytesA);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)   // 32 bytes (16 bf16) further along K = +2 in 16-byte units
                        umma_f16(tmem_d, da + uint64_t(2 * k), db + uint64_t(2 * k), (ks > 0 || k > 0) ? 1u : 0u);
                    if (CS == 1) umma_commit(empty_bar(s));   // frees the stage when these MMAs have read it
                    else umma_commit_mc(empty_bar(s), kMcMask);   // ... in every CTA of the cluster
                }
                umma_commit(tmem_full_bar(acc));              // accumulator complete
            }
        }
    } else {
        const int wq = warp & 3;                              // TMEM lane quarter this warp may read
        uint32_t n_acc = 0;
        for (int grp = first_group; grp < n_groups; grp += group_stride, ++n_acc) {
            const int acc = int(n_acc & 1u);
            mbar_wait(tmem_full_bar(acc), (n_acc >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m = (grp * CS + crank) * kBlockM + wq * 32 + lane;
            const int pos = m % kPadImage;
            const bool live = m < m_total;
            const bool pad = (pos / kPadDim == kPadDim - 1) || (pos % kPadDim == kPadDim - 1);
            uint4* orow = reinterpret_cast<uint4*>(out + size_t(m) * kChannels);
            const uint4* rrow = residual ? reinterpret_cast<const uint4*>(residual + size_t(m) * kChannels) : nullptr;
#pragma unroll 1
            for (int cc = 0; cc < kBlockN / 32; ++cc) {
                uint32_t v[32];
                tmem_ld32(tmem_base + (uint32_t(wq * 32) << 16) + uint32_t(acc) * kTmemCols + uint32_t(cc * 32), v);
                if (!live) continue;
                uint32_t packed[16];
                if (pad) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) packed[j] = 0u;
                } else {
                    uint4 r4[4];
                    if (rrow) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) r4[q] = rrow[cc * 4 + q];
                    }
                    const uint32_t* rw = reinterpret_cast<const uint32_t*>(r4);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a = __uint_as_float(v[2 * j]) + bias[cc * 32 + 2 * j];
                        float b = __uint_as_float(v[2 * j + 1]) + bias[cc * 32 + 2 * j + 1];
                        if (rrow) { a += bf16_lo(rw[j]); b += bf16_hi(rw[j]); }
                        if (relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
                        packed[j] = pack_bf16(a, b);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    orow[cc * 4 + q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
            // this warp is done reading the accumulator: hand it back to the MMA issuer
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar(acc)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CS > 1) cluster_sync_all();                  // nobody leaves while a peer may still signal its barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kTmemCols) : "memory");
    }
}


// ---- 2-SM variant (tcgen05 cta_group::2) ---------------------------------------------------------------------
// The two CTAs of a cluster (one TPC) execute ONE UMMA of M = 256: each CTA supplies the A rows of its own
// 128-row tile and HALF of the weight tile (128 of the 256 output channels), the tensor cores of both SMs read
// the two halves from both shared memories, and each SM accumulates its own 128 rows in its own TMEM.  Per CTA
// and pipeline step that is 16 KB A + 16 KB B of TMA traffic and shared-memory operand reads (the 1-SM kernel:
// 16 + 32 KB fetched/received and read) — shared-memory bandwidth is what held the 1-SM kernel at ~80 % of the
// UMMA issue floor.  Only the leader CTA (cluster rank 0) issues MMAs; both CTAs' TMA loads complete on the
// LEADER's full barrier; the leader's commits are multicast to both CTAs' empty / accumulator-full barriers;
// both CTAs' epilogue warps hand the accumulator back on the leader's barrier.
constexpr int kStages2 = 6;
constexpr uint32_t kBytesBHalf = (kBlockN / 2) * kBlockK * 2;     // 16 KB
constexpr uint32_t kBytesStage2 = kBytesA + kBytesBHalf;           // 32 KB
constexpr uint32_t kSmemBytes2 = kStages2 * kBytesStage2 + 256 + 1024;
constexpr uint32_t kInstrDesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(kBlockN >> 3) << 17) | (uint32_t(256 >> 4) << 24);

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kInstrDesc2), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

__global__ void __launch_bounds__(192, 1)
k_conv3x3_tc2(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
              const float* __restrict__ bias, const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
              int m_total, int n_tiles, int steps_per_tap, int relu) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + kStages2 * kBytesStage2;
    auto full_bar = [&](int s) { return bar_base + 8u * uint32_t(s); };
    auto empty_bar = [&](int s) { return bar_base + 8u * uint32_t(kStages2 + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * uint32_t(2 * kStages2 + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * uint32_t(2 * kStages2 + 2 + a); };
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + kStages2 * kBytesStage2 + 8 * (2 * kStages2 + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int steps = kTaps * steps_per_tap;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages2; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 8); }   // 4 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(2 * kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                              // both CTAs' barriers and TMEM exist before anyone signals a peer
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_holder;
    const int crank = int(cluster_ctarank());
    const bool leader = crank == 0;
    const int n_groups = (n_tiles + 1) / 2;
    const int first_group = int(blockIdx.x) / 2, group_stride = int(gridDim.x) / 2;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int grp = first_group; grp < n_groups; grp += group_stride) {
                const int m0 = (grp * 2 + crank) * kBlockM;
                for (int ks = 0; ks < steps; ++ks, ++it) {
                    const int s = int(it % kStages2);
                    const uint32_t ph = (it / kStages2) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);                       // freed in both CTAs by the leader's multicast commit
                    if (leader) mbar_expect_tx(full_bar(s), 2 * kBytesStage2);   // both CTAs' bytes complete on the leader's barrier
                    const uint32_t lead_full = leader ? full_bar(s) : mapa_u32(full_bar(s), 0);
                    const int tap = ks / steps_per_tap, kc = ks - tap * steps_per_tap;
                    const int shift = (tap / 3 - 1) * kPadDim + (tap % 3 - 1);
                    const uint32_t a_dst = smem_base + uint32_t(s) * kBytesStage2;
                    tma_load_2d_2sm(a_dst, &map_x, lead_full, kc * kBlockK, m0 + shift);
                    tma_load_2d_2sm(a_dst + kBytesA, &map_w, lead_full, kc * kBlockK, tap * kBlockN + crank * (kBlockN / 2));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            uint32_t it = 0, n_acc = 0;
            for (int grp = first_group; grp < n_groups; grp += group_stride, ++n_acc) {
                const int acc = int(n_acc & 1u);
                mbar_wait(tmem_empty_bar(acc), ((n_acc >> 1) & 1u) ^ 1u);  // both CTAs' epilogues have drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + uint32_t(acc) * kTmemCols;
                for (int ks = 0; ks < steps; ++ks, ++it) {
                    const int s = int(it % kStages2);
                    const uint32_t ph = (it / kStages2) & 1u;
                    mbar_wait(full_bar(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_base + uint32_t(s) * kBytesStage2;
                    const uint64_t da = umma_desc(a_addr), db = umma_desc(a_addr + kBytesA);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)
                        umma_f16_2sm(tmem_d, da + uint64_t(2 * k), db + uint64_t(2 * k), (ks > 0 || k > 0) ? 1u : 0u);
                    umma_commit_2sm(empty_bar(s), uint16_t(3));            // frees the stage in both CTAs
                }
                umma_commit_2sm(tmem_full_bar(acc), uint16_t(3));          // accumulator complete, in both CTAs
            }
        }
    } else {
        const int wq = warp & 3;
        uint32_t n_acc = 0;
        for (int grp = first_group; grp < n_groups; grp += group_stride, ++n_acc) {
            const int acc = int(n_acc & 1u);
            mbar_wait(tmem_full_bar(acc), (n_acc >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m = (grp * 2 + crank) * kBlockM + wq * 32 + lane;
            const int pos = m % kPadImage;
            const bool live = m < m_total;
            const bool pad = (pos / kPadDim == kPadDim - 1) || (pos % kPadDim == kPadDim - 1);
            uint4* orow = reinterpret_cast<uint4*>(out + size_t(m) * kChannels);
            const uint4* rrow = residual ? reinterpret_cast<const uint4*>(residual + size_t(m) * kChannels) : nullptr;
#pragma unroll 1
            for (int cc = 0; cc < kBlockN / 32; ++cc) {
                uint32_t v[32];
                tmem_ld32(tmem_base + (uint32_t(wq * 32) << 16) + uint32_t(acc) * kTmemCols + uint32_t(cc * 32), v);
                if (!live) continue;
                uint32_t packed[16];
                if (pad) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) packed[j] = 0u;
                } else {
                    uint4 r4[4];
                    if (rrow) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) r4[q] = rrow[cc * 4 + q];
                    }
                    const uint32_t* rw = reinterpret_cast<const uint32_t*>(r4);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a = __uint_as_float(v[2 * j]) + bias[cc * 32 + 2 * j];
                        float b = __uint_as_float(v[2 * j + 1]) + bias[cc * 32 + 2 * j + 1];
                        if (rrow) { a += bf16_lo(rw[j]); b += bf16_hi(rw[j]); }
                        if (relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
                        packed[j] = pack_bf16(a, b);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    orow[cc * 4 + q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {                                               // hand the accumulator back — on the LEADER's barrier
                if (leader) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar(acc)) : "memory");
                else mbar_arrive_cluster(mapa_u32(tmem_empty_bar(acc), 0));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kTmemCols) : "memory");
    }
}


// ---- 2-SM variant with A reuse across the three horizontal taps ------------------------------------------------
// The taps (dy, -1), (dy, 0), (dy, +1) read row windows of X that are one row apart, so one 136-row block
// [m0 + dy*21 - 1, +136) (17 swizzle atoms of 8 rows) serves all three: the UMMA descriptors of the three taps
// start 0, 128 and 256 bytes into the block.  The 128B swizzle is a function of the absolute shared-memory address
// (TMA wrote the block with it), so a start address one or two rows into an atom needs nothing else — measured:
// with the descriptor's base-offset field left 0 the results are bit-identical to the plain kernel's, with the
// field set to the row phase they are wrong.  A pipeline stage is one (dy, K-chunk): 17 KB of A + 3 x 16 KB of B
// per CTA and 12 UMMAs — 65 KB of TMA traffic where the plain 2-SM kernel moves 96 KB.
constexpr int kStages3 = 3;
constexpr int kRowsA3 = 136;
constexpr uint32_t kBytesA3 = kRowsA3 * kBlockK * 2;               // 17 408 (a multiple of 1024)
constexpr uint32_t kBytesStage3 = kBytesA3 + 3 * kBytesBHalf;      // 66 560
constexpr uint32_t kSmemBytes3 = kStages3 * kBytesStage3 + 256 + 1024;


__global__ void __launch_bounds__(192, 1)
k_conv3x3_tc3(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
              const float* __restrict__ bias, const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
              int m_total, int n_tiles, int steps_per_tap, int relu) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + kStages3 * kBytesStage3;
    auto full_bar = [&](int s) { return bar_base + 8u * uint32_t(s); };
    auto empty_bar = [&](int s) { return bar_base + 8u * uint32_t(kStages3 + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * uint32_t(2 * kStages3 + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * uint32_t(2 * kStages3 + 2 + a); };
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + kStages3 * kBytesStage3 + 8 * (2 * kStages3 + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int steps = 3 * steps_per_tap;              // (dy, K-chunk) super-steps

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages3; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(2 * kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_holder;
    const int crank = int(cluster_ctarank());
    const bool leader = crank == 0;
    const int n_groups = (n_tiles + 1) / 2;
    const int first_group = int(blockIdx.x) / 2, group_stride = int(gridDim.x) / 2;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int grp = first_group; grp < n_groups; grp += group_stride) {
                const int m0 = (grp * 2 + crank) * kBlockM;
                for (int ks = 0; ks < steps; ++ks, ++it) {
                    const int s = int(it % kStages3);
                    const uint32_t ph = (it / kStages3) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    if (leader) mbar_expect_tx(full_bar(s), 2 * kBytesStage3);
                    const uint32_t lead_full = leader ? full_bar(s) : mapa_u32(full_bar(s), 0);
                    const int dyi = ks / steps_per_tap, kc = ks - dyi * steps_per_tap;       // dyi = dy + 1
                    const uint32_t a_dst = smem_base + uint32_t(s) * kBytesStage3;
                    tma_load_2d_2sm(a_dst, &map_x, lead_full, kc * kBlockK, m0 + (dyi - 1) * kPadDim - 1);
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx)
                        tma_load_2d_2sm(a_dst + kBytesA3 + uint32_t(dx) * kBytesBHalf, &map_w, lead_full, kc * kBlockK,
                                        (dyi * 3 + dx) * kBlockN + crank * (kBlockN / 2));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            uint32_t it = 0, n_acc = 0;
            for (int grp = first_group; grp < n_groups; grp += group_stride, ++n_acc) {
                const int acc = int(n_acc & 1u);
                mbar_wait(tmem_empty_bar(acc), ((n_acc >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + uint32_t(acc) * kTmemCols;
                for (int ks = 0; ks < steps; ++ks, ++it) {
                    const int s = int(it % kStages3);
                    const uint32_t ph = (it / kStages3) & 1u;
                    mbar_wait(full_bar(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_base + uint32_t(s) * kBytesStage3;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const uint64_t da = umma_desc(a_addr + uint32_t(dx) * 128u);     // one row further into the block
                        const uint64_t db = umma_desc(a_addr + kBytesA3 + uint32_t(dx) * kBytesBHalf);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_f16_2sm(tmem_d, da + uint64_t(2 * k), db + uint64_t(2 * k), (ks > 0 || dx > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit_2sm(empty_bar(s), uint16_t(3));
                }
                umma_commit_2sm(tmem_full_bar(acc), uint16_t(3));
            }
        }
    } else {
        const int wq = warp & 3;
        uint32_t n_acc = 0;
        for (int grp = first_group; grp < n_groups; grp += group_stride, ++n_acc) {
            const int acc = int(n_acc & 1u);
            mbar_wait(tmem_full_bar(acc), (n_acc >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m = (grp * 2 + crank) * kBlockM + wq * 32 + lane;
            const int pos = m % kPadImage;
            const bool live = m < m_total;
            const bool pad = (pos / kPadDim == kPadDim - 1) || (pos % kPadDim == kPadDim - 1);
            uint4* orow = reinterpret_cast<uint4*>(out + size_t(m) * kChannels);
            const uint4* rrow = residual ? reinterpret_cast<const uint4*>(residual + size_t(m) * kChannels) : nullptr;
#pragma unroll 1
            for (int cc = 0; cc < kBlockN / 32; ++cc) {
                uint32_t v[32];
                tmem_ld32(tmem_base + (uint32_t(wq * 32) << 16) + uint32_t(acc) * kTmemCols + uint32_t(cc * 32), v);
                if (!live) continue;
                uint32_t packed[16];
                if (pad) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) packed[j] = 0u;
                } else {
                    uint4 r4[4];
                    if (rrow) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) r4[q] = rrow[cc * 4 + q];
                    }
                    const uint32_t* rw = reinterpret_cast<const uint32_t*>(r4);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a = __uint_as_float(v[2 * j]) + bias[cc * 32 + 2 * j];
                        float b = __uint_as_float(v[2 * j + 1]) + bias[cc * 32 + 2 * j + 1];
                        if (rrow) { a += bf16_lo(rw[j]); b += bf16_hi(rw[j]); }
                        if (relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
                        packed[j] = pack_bf16(a, b);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    orow[cc * 4 + q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (leader) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar(acc)) : "memory");
                else mbar_arrive_cluster(mapa_u32(tmem_empty_bar(acc), 0));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kTmemCols) : "memory");
    }
}


// ---- 2-SM variant with ONE A block per K-chunk for all nine taps -----------------------------------------------
// The nine taps read row windows within [m0 - 22, m0 + 150): one 176-row block (22 swizzle atoms) per K-chunk
// serves them all — tap (dy, dx) starts (dy+1)*21 + (dx+1) rows into it.  The A blocks live in their own 2-deep
// ring (filled once per K-chunk), the weights in a 3-deep ring of (dy, K-chunk) stages (3 x 16 KB each), so the
// per-CTA TMA traffic per K-chunk is 22 + 144 KB (the per-dy A blocks of k_conv3x3_tc3: 51 + 144 KB).
constexpr int kRowsA4 = 176;
constexpr uint32_t kBytesA4 = kRowsA4 * kBlockK * 2;               // 22 528
constexpr uint32_t kBytesB4 = 3 * kBytesBHalf;                     // 49 152 per (dy, K-chunk) stage
constexpr int kStagesA4 = 2, kStagesB4 = 3;
constexpr uint32_t kSmemBytes4 = kStagesA4 * kBytesA4 + kStagesB4 * kBytesB4 + 256 + 1024;

__global__ void __launch_bounds__(192, 1)
k_conv3x3_tc4(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
              const float* __restrict__ bias, const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
              int m_total, int n_tiles, int steps_per_tap, int relu, int tile0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t b_base = smem_base + kStagesA4 * kBytesA4;
    const uint32_t bar_base = b_base + kStagesB4 * kBytesB4;
    auto a_full = [&](int s) { return bar_base + 8u * uint32_t(s); };
    auto a_empty = [&](int s) { return bar_base + 8u * uint32_t(2 + s); };
    auto b_full = [&](int s) { return bar_base + 8u * uint32_t(4 + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * uint32_t(7 + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * uint32_t(10 + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * uint32_t(12 + a); };
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + kStagesA4 * kBytesA4 + kStagesB4 * kBytesB4 + 8 * 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStagesA4; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < kStagesB4; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(2 * kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_holder;
    const int crank = int(cluster_ctarank());
    const bool leader = crank == 0;
    const int n_groups = (n_tiles + 1) / 2;
    const int first_group = int(blockIdx.x) / 2, group_stride = int(gridDim.x) / 2;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t ia = 0, ib = 0;                                   // A-block / B-stage counters across tiles
            for (int grp = first_group; grp < n_groups; grp += group_stride) {
                const int m0 = (tile0 + grp * 2 + crank) * kBlockM;
                for (int kc = 0; kc < steps_per_tap; ++kc, ++ia) {
                    const int sa = int(ia % kStagesA4);
                    mbar_wait(a_empty(sa), ((ia / kStagesA4) & 1u) ^ 1u);
                    if (leader) mbar_expect_tx(a_full(sa), 2 * kBytesA4);
                    tma_load_2d_2sm(smem_base + uint32_t(sa) * kBytesA4, &map_x, leader ? a_full(sa) : mapa_u32(a_full(sa), 0),
                                    kc * kBlockK, m0 - (kPadDim + 1));
                    for (int dyi = 0; dyi < 3; ++dyi, ++ib) {
                        const int sb = int(ib % kStagesB4);
                        mbar_wait(b_empty(sb), ((ib / kStagesB4) & 1u) ^ 1u);
                        if (leader) mbar_expect_tx(b_full(sb), 2 * kBytesB4);
                        const uint32_t lead_full = leader ? b_full(sb) : mapa_u32(b_full(sb), 0);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
                            tma_load_2d_2sm(b_base + uint32_t(sb) * kBytesB4 + uint32_t(dx) * kBytesBHalf, &map_w, lead_full, kc * kBlockK,
                                            (dyi * 3 + dx) * kBlockN + crank * (kBlockN / 2));
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            uint32_t ia = 0, ib = 0, n_acc = 0;
            for (int grp = first_group; grp < n_groups; grp += group_stride, ++n_acc) {
                const int acc = int(n_acc & 1u);
                mbar_wait(tmem_empty_bar(acc), ((n_acc >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + uint32_t(acc) * kTmemCols;
                for (int kc = 0; kc < steps_per_tap; ++kc, ++ia) {
                    const int sa = int(ia % kStagesA4);
                    mbar_wait(a_full(sa), (ia / kStagesA4) & 1u);
                    const uint32_t a_addr = smem_base + uint32_t(sa) * kBytesA4;
                    for (int dyi = 0; dyi < 3; ++dyi, ++ib) {
                        const int sb = int(ib % kStagesB4);
                        mbar_wait(b_full(sb), (ib / kStagesB4) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t bs = b_base + uint32_t(sb) * kBytesB4;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint64_t da = umma_desc(a_addr + uint32_t(dyi * kPadDim + dx) * 128u);   // rows into the block
                            const uint64_t db = umma_desc(bs + uint32_t(dx) * kBytesBHalf);
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k)
                                umma_f16_2sm(tmem_d, da + uint64_t(2 * k), db + uint64_t(2 * k), (kc > 0 || dyi > 0 || dx > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit_2sm(b_empty(sb), uint16_t(3));
                    }
                    umma_commit_2sm(a_empty(sa), uint16_t(3));         // the A block is free once its 36 UMMAs have read it
                }
                umma_commit_2sm(tmem_full_bar(acc), uint16_t(3));
            }
        }
    } else {
        const int wq = warp & 3;
        uint32_t n_acc = 0;
        for (int grp = first_group; grp < n_groups; grp += group_stride, ++n_acc) {
            const int acc = int(n_acc & 1u);
            mbar_wait(tmem_full_bar(acc), (n_acc >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m = (tile0 + grp * 2 + crank) * kBlockM + wq * 32 + lane;
            const int pos = m % kPadImage;
            const bool live = m < m_total;
            const bool pad = (pos / kPadDim == kPadDim - 1) || (pos % kPadDim == kPadDim - 1);
            uint4* orow = reinterpret_cast<uint4*>(out + size_t(m) * kChannels);
            const uint4* rrow = residual ? reinterpret_cast<const uint4*>(residual + size_t(m) * kChannels) : nullptr;
#pragma unroll 1
            for (int cc = 0; cc < kBlockN / 32; ++cc) {
                uint32_t v[32];
                tmem_ld32(tmem_base + (uint32_t(wq * 32) << 16) + uint32_t(acc) * kTmemCols + uint32_t(cc * 32), v);
                if (!live) continue;
                uint32_t packed[16];
                if (pad) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) packed[j] = 0u;
                } else {
                    uint4 r4[4];
                    if (rrow) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) r4[q] = rrow[cc * 4 + q];
                    }
                    const uint32_t* rw = reinterpret_cast<const uint32_t*>(r4);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a = __uint_as_float(v[2 * j]) + bias[cc * 32 + 2 * j];
                        float b = __uint_as_float(v[2 * j + 1]) + bias[cc * 32 + 2 * j + 1];
                        if (rrow) { a += bf16_lo(rw[j]); b += bf16_hi(rw[j]); }
                        if (relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
                        packed[j] = pack_bf16(a, b);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    orow[cc * 4 + q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (leader) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar(acc)) : "memory");
                else mbar_arrive_cluster(mapa_u32(tmem_empty_bar(acc), 0));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kTmemCols) : "memory");
    }
}



__device__ __forceinline__ void tma_load_2d_2sm_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
// barrier operand of the multicast weight loads: every destination CTA's PAIR LEADER must see the bytes.
#ifndef BK_CONV5_BAR_MODE
#define BK_CONV5_BAR_MODE 0
#endif
#if BK_CONV5_BAR_MODE == 0
#define BK_CONV5_BAR(b) ((b) & 0xFEFFFFFFu)                       /* CTA-relative offset with the pair-peer bit cleared */
#elif BK_CONV5_BAR_MODE == 1
#define BK_CONV5_BAR(b) (leader ? (b) : mapa_u32((b), lead_rank)) /* explicit address of this CTA's pair leader */
#else
#define BK_CONV5_BAR(b) (b)
#endif
// ---- the same, in clusters of FOUR: the weight stream is shared by two CTA pairs ----------------------------------
// CTAs 0/1 and 2/3 of a cluster are two MMA pairs working on different row tiles with the SAME weights: each CTA
// fetches only a quarter of a weight tile (64 output channels) and TMA-multicasts it to the CTA of the other pair
// that needs the same half (0 <-> 2, 1 <-> 3), so the per-CTA weight fetch halves (72 KB per K-chunk instead of 144).
// A weight stage is recycled when BOTH pairs' MMAs have consumed it (both leaders' commits reach all four CTAs).
// (text of the 2-CTA kernel follows)
// ---- 2-SM variant with ONE A block per K-chunk for all nine taps
// The nine taps read row windows within [m0 - 22, m0 + 150): one 176-row block (22 swizzle atoms) per K-chunk
// serves them all — tap (dy, dx) starts (dy+1)*21 + (dx+1) rows into it.  The A blocks live in their own 2-deep
// ring (filled once per K-chunk), the weights in a 3-deep ring of (dy, K-chunk) stages (3 x 16 KB each), so the
// per-CTA TMA traffic per K-chunk is 22 + 144 KB (the per-dy A blocks of k_conv3x3_tc3: 51 + 144 KB).

__global__ void __launch_bounds__(192, 1)
k_conv3x3_tc5(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
              const float* __restrict__ bias, const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
              int m_total, int n_tiles, int steps_per_tap, int relu, int tile0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t b_base = smem_base + kStagesA4 * kBytesA4;
    const uint32_t bar_base = b_base + kStagesB4 * kBytesB4;
    auto a_full = [&](int s) { return bar_base + 8u * uint32_t(s); };
    auto a_empty = [&](int s) { return bar_base + 8u * uint32_t(2 + s); };
    auto b_full = [&](int s) { return bar_base + 8u * uint32_t(4 + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * uint32_t(7 + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * uint32_t(10 + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * uint32_t(12 + a); };
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + kStagesA4 * kBytesA4 + kStagesB4 * kBytesB4 + 8 * 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStagesA4; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < kStagesB4; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 2); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(2 * kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_holder;
    const int crank = int(cluster_ctarank());        // 0..3
    const bool leader = (crank & 1) == 0;
    const uint32_t lead_rank = uint32_t(crank & ~1);
    const int half = crank & 1;                      // which half of the weight tile this CTA's pair member holds
    const int quarter = crank >> 1;                  // which quarter of that half this CTA fetches
    const uint16_t side_mask = uint16_t(0x5u << half);              // CTAs holding the same half: {0,2} or {1,3}
    const uint16_t pair_mask = uint16_t(0x3u << (crank & ~1));      // this CTA's MMA pair
    const int n_groups = (n_tiles + 3) / 4;
    const int first_group = int(blockIdx.x) / 4, group_stride = int(gridDim.x) / 4;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t ia = 0, ib = 0;                                   // A-block / B-stage counters across tiles
            for (int grp = first_group; grp < n_groups; grp += group_stride) {
                const int m0 = (tile0 + grp * 4 + crank) * kBlockM;
                for (int kc = 0; kc < steps_per_tap; ++kc, ++ia) {
                    const int sa = int(ia % kStagesA4);
                    mbar_wait(a_empty(sa), ((ia / kStagesA4) & 1u) ^ 1u);
                    if (leader) mbar_expect_tx(a_full(sa), 2 * kBytesA4);
                    tma_load_2d_2sm(smem_base + uint32_t(sa) * kBytesA4, &map_x, leader ? a_full(sa) : mapa_u32(a_full(sa), lead_rank),
                                    kc * kBlockK, m0 - (kPadDim + 1));
                    for (int dyi = 0; dyi < 3; ++dyi, ++ib) {
                        const int sb = int(ib % kStagesB4);
                        mbar_wait(b_empty(sb), ((ib / kStagesB4) & 1u) ^ 1u);
                        if (leader) mbar_expect_tx(b_full(sb), 2 * kBytesB4);
                        const uint32_t lead_full = BK_CONV5_BAR(b_full(sb));
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
                            tma_load_2d_2sm_mc(b_base + uint32_t(sb) * kBytesB4 + uint32_t(dx) * kBytesBHalf + uint32_t(quarter) * (kBytesBHalf / 2),
                                               &map_w, lead_full, kc * kBlockK,
                                               (dyi * 3 + dx) * kBlockN + half * (kBlockN / 2) + quarter * (kBlockN / 4), side_mask);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            uint32_t ia = 0, ib = 0, n_acc = 0;
            for (int grp = first_group; grp < n_groups; grp += group_stride, ++n_acc) {
                const int acc = int(n_acc & 1u);
                mbar_wait(tmem_empty_bar(acc), ((n_acc >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + uint32_t(acc) * kTmemCols;
                for (int kc = 0; kc < steps_per_tap; ++kc, ++ia) {
                    const int sa = int(ia % kStagesA4);
                    mbar_wait(a_full(sa), (ia / kStagesA4) & 1u);
                    const uint32_t a_addr = smem_base + uint32_t(sa) * kBytesA4;
                    for (int dyi = 0; dyi < 3; ++dyi, ++ib) {
                        const int sb = int(ib % kStagesB4);
                        mbar_wait(b_full(sb), (ib / kStagesB4) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t bs = b_base + uint32_t(sb) * kBytesB4;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint64_t da = umma_desc(a_addr + uint32_t(dyi * kPadDim + dx) * 128u);   // rows into the block
                            const uint64_t db = umma_desc(bs + uint32_t(dx) * kBytesBHalf);
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k)
                                umma_f16_2sm(tmem_d, da + uint64_t(2 * k), db + uint64_t(2 * k), (kc > 0 || dyi > 0 || dx > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit_2sm(b_empty(sb), uint16_t(0xF));              // all four CTAs: both pairs share the stage
                    }
                    umma_commit_2sm(a_empty(sa), pair_mask);         // the A block is free once its 36 UMMAs have read it
                }
                umma_commit_2sm(tmem_full_bar(acc), pair_mask);
            }
        }
    } else {
        const int wq = warp & 3;
        uint32_t n_acc = 0;
        for (int grp = first_group; grp < n_groups; grp += group_stride, ++n_acc) {
            const int acc = int(n_acc & 1u);
            mbar_wait(tmem_full_bar(acc), (n_acc >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m = (tile0 + grp * 4 + crank) * kBlockM + wq * 32 + lane;
            const int pos = m % kPadImage;
            const bool live = m < m_total;
            const bool pad = (pos / kPadDim == kPadDim - 1) || (pos % kPadDim == kPadDim - 1);
            uint4* orow = reinterpret_cast<uint4*>(out + size_t(m) * kChannels);
            const uint4* rrow = residual ? reinterpret_cast<const uint4*>(residual + size_t(m) * kChannels) : nullptr;
#pragma unroll 1
            for (int cc = 0; cc < kBlockN / 32; ++cc) {
                uint32_t v[32];
                tmem_ld32(tmem_base + (uint32_t(wq * 32) << 16) + uint32_t(acc) * kTmemCols + uint32_t(cc * 32), v);
                if (!live) continue;
                uint32_t packed[16];
                if (pad) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) packed[j] = 0u;
                } else {
                    uint4 r4[4];
                    if (rrow) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) r4[q] = rrow[cc * 4 + q];
                    }
                    const uint32_t* rw = reinterpret_cast<const uint32_t*>(r4);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a = __uint_as_float(v[2 * j]) + bias[cc * 32 + 2 * j];
                        float b = __uint_as_float(v[2 * j + 1]) + bias[cc * 32 + 2 * j + 1];
                        if (rrow) { a += bf16_lo(rw[j]); b += bf16_hi(rw[j]); }
                        if (relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
                        packed[j] = pack_bf16(a, b);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    orow[cc * 4 + q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (leader) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar(acc)) : "memory");
                else mbar_arrive_cluster(mapa_u32(tmem_empty_bar(acc), lead_rank));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kTmemCols) : "memory");
    }
}



typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows][cols] bf16 matrix, box = 64 channels x box_rows rows, 128B swizzle, zero fill outside
int make_map(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return bk_fail(BK_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {kBlockK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return bk_fail(BK_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")");
    return BK_OK;
}

}  // namespace
#endif  // BK_WARP_EMU

static int conv_launch(const void* dev_x, const void* dev_w, const float* dev_bias, const void* dev_residual, void* dev_y,
                       int batch, int in_channels, int relu, void* cuda_stream) {
#ifdef BK_WARP_EMU
    (void)dev_x; (void)dev_w; (void)dev_bias; (void)dev_residual; (void)dev_y; (void)batch; (void)in_channels; (void)relu; (void)cuda_stream;
    return bk_fail(BK_ERR_STATE, "bk_conv3x3_bf16: tensor-core kernel, not available in the CPU emulator build");
#else
    if (!dev_x || !dev_w || !dev_bias || !dev_y || batch <= 0) return bk_fail(BK_ERR_INVALID_ARG, "bk_conv3x3_bf16: bad argument");
    if (in_channels <= 0 || in_channels > kChannels || in_channels % kBlockK) return bk_fail(BK_ERR_INVALID_ARG, "bk_conv3x3_bf16: in_channels must be 64, 128, 192 or 256");
    static int n_sm = 0;
    static int cluster = 2;
    // BK_CONV_QUAD=1: clusters of four, the weight stream multicast across the two CTA pairs.  Correct, and 6.5 % faster
    // per SM, but only 33 clusters of four are co-resident on the 148 SMs (a cluster must fit inside a GPC), so 16 SMs
    // idle: 0.321 ms per convolution against 0.305 ms for the pair kernel.  With CTA pairs on the left-over SMs from a second
    // stream (the hybrid below; BK_CONV_QUAD_ONLY=1 disables it) the launch ties with the pair kernel (0.306 ms).  An option.
    static int quad = 0;
    static int a_reuse9 = 1;        // one A block per K-chunk for all nine taps: 0.305 ms per convolution at batch 1024 (BK_CONV_AREUSE9=0: per-dy blocks, 0.311)
    static int a_reuse = 1;         // the 2-SM kernel with one A block per (dy, K-chunk): 0.321 ms per convolution at batch 1024 (BK_CONV_AREUSE=0: 0.344)
    static int two_sm = 1;          // the cta_group::2 kernel (0.348 ms against 0.370 ms per convolution at batch 1024); BK_CONV_2SM=0 selects the 1-SM one
    if (!n_sm) {
        BK_CUDA(cudaFuncSetAttribute(k_conv3x3_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes2)));
        if (const char* e = getenv("BK_CONV_2SM")) two_sm = atoi(e);
        BK_CUDA(cudaFuncSetAttribute(k_conv3x3_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes3)));
        if (const char* e = getenv("BK_CONV_AREUSE")) a_reuse = atoi(e);
        BK_CUDA(cudaFuncSetAttribute(k_conv3x3_tc4, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes4)));
        BK_CUDA(cudaFuncSetAttribute(k_conv3x3_tc5, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes4)));
        if (const char* e = getenv("BK_CONV_QUAD")) quad = atoi(e);
        if (quad) { a_reuse9 = 1; }
        if (const char* e = getenv("BK_CONV_AREUSE9")) a_reuse9 = quad ? 1 : atoi(e);
        if (a_reuse9) a_reuse = 1;
        if (a_reuse) two_sm = 1;
        BK_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
        BK_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
        BK_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
        if (const char* e = getenv("BK_CONV_CLUSTER")) cluster = atoi(e);
        if (cluster != 1 && cluster != 2 && cluster != 4) cluster = 2;
        int dev = 0;
        BK_CUDA(cudaGetDevice(&dev));
        BK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int m_total = batch * kPadImage;
    CUtensorMap map_x, map_w;
    int rc = make_map(&map_x, dev_x, uint64_t(in_channels), uint64_t(m_total), a_reuse9 ? uint32_t(kRowsA4) : (a_reuse ? uint32_t(kRowsA3) : uint32_t(kBlockM)));
    if (rc) return rc;
    if (two_sm) cluster = 2;
    if (quad) cluster = 4;
    rc = make_map(&map_w, dev_w, uint64_t(in_channels), uint64_t(kTaps) * kBlockN, uint32_t(kBlockN / cluster));
    if (rc) return rc;
    const int tiles = (m_total + kBlockM - 1) / kBlockM;
    const int groups = (tiles + cluster - 1) / cluster;
    int n_clusters = n_sm / cluster;
    if (groups < n_clusters) n_clusters = groups;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(n_clusters * cluster));
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = a_reuse9 ? kSmemBytes4 : a_reuse ? kSmemBytes3 : (two_sm ? kSmemBytes2 : kSmemBytes);
    cfg.stream = static_cast<cudaStream_t>(cuda_stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = unsigned(cluster);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (quad) {
        // clusters of four must fit inside a GPC, so fewer than n_sm / 4 may be co-resident: a persistent grid larger
        // than that would run its last clusters after the others
        static int max_quads = 0;
        if (!max_quads) {
            cfg.gridDim = dim3(unsigned((n_sm / 4) * 4));
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, k_conv3x3_tc5, &cfg) == cudaSuccess && n > 0) max_quads = n;
            else max_quads = n_sm / 4;
            if (getenv("BK_CONV_DEBUG")) fprintf(stderr, "bk_conv: %d clusters of 4 can be co-resident on %d SMs\n", max_quads, n_sm);
        }
        if (n_clusters > max_quads) n_clusters = max_quads;
        cfg.gridDim = dim3(unsigned(n_clusters * cluster));
    }
    const __nv_bfloat16* res = static_cast<const __nv_bfloat16*>(dev_residual);
    __nv_bfloat16* y = static_cast<__nv_bfloat16*>(dev_y);
    const int spt = in_channels / kBlockK;
    if (quad) {
        // Hybrid: clusters of four on the SMs that can host them, CTA pairs (second stream, concurrently) on the SMs
        // left over in each GPC; the tiles are split in proportion to the two grids' per-SM speeds.
        static cudaStream_t side = nullptr;
        static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
        if (!side) {
            BK_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
            BK_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
            BK_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        }
        const int quad_sms = n_clusters * 4;
        const int rest_pairs = (n_sm - quad_sms) / 2;
        int tiles_pair = 0;
        if (rest_pairs > 0 && tiles > 8 * n_sm && !getenv("BK_CONV_QUAD_ONLY")) {
            double share = (2.0 * rest_pairs) / (2.0 * rest_pairs + 1.065 * quad_sms);
            if (const char* e = getenv("BK_CONV_PAIR_SHARE")) share = atof(e);       // probes
            tiles_pair = (int(share * tiles) / 2) * 2;
        }
        const int tiles_quad = tiles - tiles_pair;
        if (tiles_pair > 0) BK_CUDA(cudaEventRecord(ev_fork, cfg.stream));      // fork BEFORE the first kernel is enqueued
        BK_CUDA(cudaLaunchKernelEx(&cfg, k_conv3x3_tc5, map_x, map_w, dev_bias, res, y, m_total, tiles_quad, spt, relu, 0));
        if (tiles_pair > 0) {
            CUtensorMap map_w2;                                                // weight map with 128-row boxes for the pair kernel
            rc = make_map(&map_w2, dev_w, uint64_t(in_channels), uint64_t(kTaps) * kBlockN, uint32_t(kBlockN / 2));
            if (rc) return rc;
            BK_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
            cudaLaunchConfig_t cfg2 = cfg;
            cudaLaunchAttribute attr2[1];
            attr2[0] = attr[0];
            attr2[0].val.clusterDim.x = 2;
            cfg2.attrs = attr2;
            cfg2.gridDim = dim3(unsigned(rest_pairs * 2));
            cfg2.stream = side;
            BK_CUDA(cudaLaunchKernelEx(&cfg2, k_conv3x3_tc4, map_x, map_w2, dev_bias, res, y, m_total, tiles_pair, spt, relu, tiles_quad));
            BK_CUDA(cudaEventRecord(ev_join, side));
            BK_CUDA(cudaStreamWaitEvent(cfg.stream, ev_join, 0));
        }
    }
    else if (a_reuse9) BK_CUDA(cudaLaunchKernelEx(&cfg, k_conv3x3_tc4, map_x, map_w, dev_bias, res, y, m_total, tiles, spt, relu, 0));
    else if (a_reuse) BK_CUDA(cudaLaunchKernelEx(&cfg, k_conv3x3_tc3, map_x, map_w, dev_bias, res, y, m_total, tiles, spt, relu));
    else if (two_sm) BK_CUDA(cudaLaunchKernelEx(&cfg, k_conv3x3_tc2, map_x, map_w, dev_bias, res, y, m_total, tiles, spt, relu));
    else if (cluster == 1) BK_CUDA(cudaLaunchKernelEx(&cfg, k_conv3x3_tc<1>, map_x, map_w, dev_bias, res, y, m_total, tiles, spt, relu));
    else if (cluster == 2) BK_CUDA(cudaLaunchKernelEx(&cfg, k_conv3x3_tc<2>, map_x, map_w, dev_bias, res, y, m_total, tiles, spt, relu));
    else BK_CUDA(cudaLaunchKernelEx(&cfg, k_conv3x3_tc<4>, map_x, map_w, dev_bias, res, y, m_total, tiles, spt, relu));
    BK_CUDA(cudaGetLastError());
    return BK_OK;
#endif
}

extern "C" int bk_conv3x3_bf16(const void* dev_x, const void* dev_w, const float* dev_bias, const void* dev_residual,
                               void* dev_y, int batch, int relu, void* cuda_stream) {
    return conv_launch(dev_x, dev_w, dev_bias, dev_residual, dev_y, batch, 256, relu, cuda_stream);
}

extern "C" int bk_conv3x3_bf16_in(const void* dev_x, const void* dev_w, const float* dev_bias, void* dev_y, int batch,
                                  int in_channels, int relu, void* cuda_stream) {
    return conv_launch(dev_x, dev_w, dev_bias, nullptr, dev_y, batch, in_channels, relu, cuda_stream);
}
