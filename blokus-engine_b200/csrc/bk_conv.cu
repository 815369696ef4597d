// bk_conv.cu — hand-written sm_100a tensor-core kernel for the one dense contraction on the path: the
// 3x3 convolutions of the reference's policy/value ResNet trunk (model/resnet.py:8-26,51-52;
// SURVEY.md §8f row f2).  tcgen05.mma (cta_group::2) with the accumulator in TMEM, operands staged by TMA into
// 128B-swizzled shared memory, BatchNorm folded into weights/bias, bias + residual + ReLU fused into the
// epilogue that reads TMEM with tcgen05.ld.
//
// Formulation.  Activations live in HBM as a zero-padded NHWC matrix X[m][c], bf16, with
// m = image*441 + row*21 + col (row, col in 0..20; row 20 and col 20 are zero padding shared by neighbours),
// so a 3x3 tap (dy, dx) is a pure row shift of dy*21 + dx and the convolution is nine shifted GEMMs
//     Y[m][n] = sum_tap sum_k X[m + dy*21 + dx][k] * W[tap][n][k]
// whose A tiles are plain 2-D TMA boxes (out-of-range rows are zero-filled by TMA).
//
// One kernel, k_conv3x3_tc4 (the earlier generations — 1-SM clusters with multicast weights, plain 2-SM, per-dy A
// blocks, clusters of four — are in the git history and described in DESIGN.md §3; this is the one that won):
//   * CTA pairs (one TPC) issue ONE tcgen05.mma.cta_group::2 of M = 256, N = 256, K = 16: each CTA supplies the A
//     rows of its own 128-row tile and HALF of the weight tile (128 of the 256 output channels); the tensor cores of
//     both SMs read the two halves from both shared memories; each SM accumulates its own rows in its own TMEM.
//     Only the leader CTA issues MMAs; both CTAs' TMA loads complete on the LEADER's full barrier; the leader's
//     commits are multicast to both CTAs' empty / accumulator-full barriers.
//   * ONE 176-row A block per K-chunk of 64 channels serves all nine taps (their windows lie in [m0 - 22, m0 + 150));
//     tap (dy, dx) starts (dy+1)*21 + (dx+1) rows into the block — the 128B swizzle is a function of the absolute
//     shared-memory address (TMA wrote the block with it), so a descriptor may start mid-atom.  A blocks: 2-deep
//     ring; weights: 3-deep ring of (dy, K-chunk) stages of 3 x 16 KB.
//   * persistent CTAs, accumulator double-buffered in TMEM (2 x 256 columns): the epilogue of tile i overlaps the
//     TMA/MMA main loop of tile i + 1.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane of the
// pair's leader CTA), warps 2..5 = epilogue (each owns the 32 TMEM lanes of its warp-in-warpgroup rank).
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
constexpr int kTaps = 9;
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

// ---- cta_group::2 primitives ---------------------------------------------------------------------------------
constexpr uint32_t kBytesBHalf = (kBlockN / 2) * kBlockK * 2;     // 16 KB
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

// ---- the kernel ----------------------------------------------------------------------------------------------
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

int bk_conv_launch(const void* dev_x, const void* dev_w, const float* dev_bias, const void* dev_residual, void* dev_y,
                   int batch, int in_channels, int relu, void* cuda_stream) {
#ifdef BK_WARP_EMU
    (void)dev_x; (void)dev_w; (void)dev_bias; (void)dev_residual; (void)dev_y; (void)batch; (void)in_channels; (void)relu; (void)cuda_stream;
    return bk_fail(BK_ERR_STATE, "bk_conv3x3_bf16: tensor-core kernel, not available in the CPU emulator build");
#else
    if (!dev_x || !dev_w || !dev_bias || !dev_y || batch <= 0) return bk_fail(BK_ERR_INVALID_ARG, "bk_conv3x3_bf16: bad argument");
    if (in_channels <= 0 || in_channels > kChannels || in_channels % kBlockK) return bk_fail(BK_ERR_INVALID_ARG, "bk_conv3x3_bf16: in_channels must be 64, 128, 192 or 256");
    int dev = 0;
    BK_CUDA(cudaGetDevice(&dev));
    static int n_sm_of[64] = {0};            // per device: SM count, 0 = kernel attributes not yet set on that device
    if (dev < 0 || dev >= 64) return bk_fail(BK_ERR_INVALID_ARG, "bk_conv3x3_bf16: device index out of range");
    if (!n_sm_of[dev]) {
        BK_CUDA(cudaFuncSetAttribute(k_conv3x3_tc4, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes4)));
        int n = 0;
        BK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        n_sm_of[dev] = n;
    }
    const int n_sm = n_sm_of[dev];
    const int m_total = batch * kPadImage;
    CUtensorMap map_x, map_w;
    int rc = make_map(&map_x, dev_x, uint64_t(in_channels), uint64_t(m_total), uint32_t(kRowsA4));
    if (rc) return rc;
    rc = make_map(&map_w, dev_w, uint64_t(in_channels), uint64_t(kTaps) * kBlockN, uint32_t(kBlockN / 2));
    if (rc) return rc;
    const int tiles = (m_total + kBlockM - 1) / kBlockM;
    const int groups = (tiles + 1) / 2;
    int n_pairs = n_sm / 2;                       // persistent: one CTA pair per TPC
    if (groups < n_pairs) n_pairs = groups;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(n_pairs * 2));
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = kSmemBytes4;
    cfg.stream = static_cast<cudaStream_t>(cuda_stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    BK_CUDA(cudaLaunchKernelEx(&cfg, k_conv3x3_tc4, map_x, map_w, dev_bias, static_cast<const __nv_bfloat16*>(dev_residual),
                               static_cast<__nv_bfloat16*>(dev_y), m_total, tiles, in_channels / kBlockK, relu, 0));
    BK_CUDA(cudaGetLastError());
    return BK_OK;
#endif
}

extern "C" int bk_conv3x3_bf16(const void* dev_x, const void* dev_w, const float* dev_bias, const void* dev_residual,
                               void* dev_y, int batch, int relu, void* cuda_stream) {
    return bk_conv_launch(dev_x, dev_w, dev_bias, dev_residual, dev_y, batch, 256, relu, cuda_stream);
}

extern "C" int bk_conv3x3_bf16_in(const void* dev_x, const void* dev_w, const float* dev_bias, void* dev_y, int batch,
                                  int in_channels, int relu, void* cuda_stream) {
    return bk_conv_launch(dev_x, dev_w, dev_bias, nullptr, dev_y, batch, in_channels, relu, cuda_stream);
}
