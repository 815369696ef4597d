// bk_eval_kernels.cuh — everything of the leaf evaluator's round that is NOT a 3x3 convolution
// (model/resnet.py:69-94, model/training.py:43-67), written so that the CPU warp emulator of the tests
// executes the same source:
//
//   kb_planes_nhwc   Game::get_board_state (game.rs:283-311) of a position, written straight into the
//                    convolution kernel's input layout: zero-padded NHWC bf16 [row*441 + r*21 + c][64]
//                    (channels 0..4 = the five planes, 5..63 stay zero), no float planes in between
//   kb_pack_planes   the same layout from caller-supplied float planes [R][5][20][20]
//   kb_heads         policy head (1x1 conv -> BN -> ReLU -> masked softmax x mask) and value head
//                    (1x1 conv -> BN -> ReLU -> Linear(400, 4) -> tanh -> softmax) of one position,
//                    reading the trunk's final bf16 activations once
#pragma once
#include "bk_game.cuh"

#define BK_PAD_DIM 21
#define BK_PAD_IMAGE 441
#define BK_EVAL_IN_CH 64      // input channels of the first convolution as the kernel sees them (5 real ones)
#define BK_EVAL_CH 256

__device__ __forceinline__ uint32_t bk_f32_to_bf16_bits(float f) {      // round to nearest even (finite inputs)
    const uint32_t u = __float_as_uint(f);
    return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}
__device__ __forceinline__ float bk_bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bk_bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

#define BK_BF16_ONE 0x3F80u

// board (r, c) of mover-frame cell (j, k) for seat `cur`: new[j][k] = old[k][19-j], applied cur times
__device__ __forceinline__ void bk_unrotate(int j, int k, int cur, int& r, int& c) {
    if (cur == 0) { r = j; c = k; }
    else if (cur == 1) { r = k; c = 19 - j; }
    else if (cur == 2) { r = 19 - j; c = 19 - k; }
    else { r = 19 - k; c = j; }
}

// One position: 400 16-byte stores (channels 0..7 of each cell; 5..7 are zeros).
__device__ __forceinline__ void kb_planes_nhwc(const BkState* __restrict__ s, uint4* __restrict__ x64_row0, int tid, int nthreads) {
    const int cur = int(s->meta & 3u);
    for (int e = tid; e < 400; e += nthreads) {
        const int j = e / 20, k = e % 20;
        int r, c;
        bk_unrotate(j, k, cur, r, c);
        const uint4 own = reinterpret_cast<const uint4*>(s->own)[r];
        const uint32_t o[4] = {own.x, own.y, own.z, own.w};
        uint32_t b[5];
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) b[pl] = ((o[(pl + cur) & 3] >> c) & 1u) ? BK_BF16_ONE : 0u;
        b[4] = ((s->legal[r] >> c) & 1u) ? BK_BF16_ONE : 0u;
        x64_row0[size_t(j * BK_PAD_DIM + k) * (BK_EVAL_IN_CH / 8)] = make_uint4(b[0] | (b[1] << 16), b[2] | (b[3] << 16), b[4], 0u);
    }
}

__device__ __forceinline__ void kb_pack_planes(const float* __restrict__ planes /* [5][20][20] */, uint4* __restrict__ x64_row0,
                                               int tid, int nthreads) {
    for (int e = tid; e < 400; e += nthreads) {
        uint32_t b[5];
#pragma unroll
        for (int pl = 0; pl < 5; ++pl) b[pl] = bk_f32_to_bf16_bits(planes[pl * 400 + e]);
        x64_row0[size_t((e / 20) * BK_PAD_DIM + e % 20) * (BK_EVAL_IN_CH / 8)] = make_uint4(b[0] | (b[1] << 16), b[2] | (b[3] << 16), b[4], 0u);
    }
}

struct BkHeadParams {
    const float* head_w;      // [2][256]: policy 1x1 convolution, value 1x1 convolution
    const float* head_affine; // [4]: policy scale, policy shift, value scale, value shift (conv bias + eval BatchNorm folded)
    const float* lin_w;       // [4][400]
    const float* lin_b;       // [4]
};

// One CTA per position.  act = the trunk's output for this position's image, padded NHWC bf16 [441][256];
// x64 = the image's input (channel 4 = legal mask).  sh = 800 floats of shared memory + 16 floats of scratch.
// Outputs: policy[400] (mover frame, zero on illegal tiles), value[4]; optional logits[400] / vtanh[4]
// (the pre-softmax quantities, for parity tests against the fp32 reference).
__device__ __forceinline__ void kb_heads(const uint4* __restrict__ act, const uint4* __restrict__ x64, const BkHeadParams& hp,
                                         float* __restrict__ policy, float* __restrict__ value, float* __restrict__ logits,
                                         float* __restrict__ vtanh, float* sh, int tid, int nthreads) {
    float* s_logit = sh;            // [400]
    float* s_hv = sh + 400;         // [400]
    float* s_red = sh + 800;        // [16]
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    float wp[8], wv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { wp[i] = hp.head_w[lane * 8 + i]; wv[i] = hp.head_w[BK_EVAL_CH + lane * 8 + i]; }
    const float ps = hp.head_affine[0], pb = hp.head_affine[1], vs = hp.head_affine[2], vb = hp.head_affine[3];
    for (int p = warp; p < 400; p += nwarps) {
        const int m = (p / 20) * BK_PAD_DIM + p % 20;
        const uint4 a = act[size_t(m) * (BK_EVAL_CH / 8) + lane];           // 8 channels of this cell
        const float x[8] = {bk_bf16_lo(a.x), bk_bf16_hi(a.x), bk_bf16_lo(a.y), bk_bf16_hi(a.y),
                            bk_bf16_lo(a.z), bk_bf16_hi(a.z), bk_bf16_lo(a.w), bk_bf16_hi(a.w)};
        float dp = 0.0f, dv = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { dp += x[i] * wp[i]; dv += x[i] * wv[i]; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { dp += __shfl_xor_sync(BK_FULL, dp, d); dv += __shfl_xor_sync(BK_FULL, dv, d); }
        if (lane == 0) {
            const float lp = dp * ps + pb, lv = dv * vs + vb;
            s_logit[p] = lp > 0.0f ? lp : 0.0f;
            s_hv[p] = lv > 0.0f ? lv : 0.0f;
        }
    }
    __syncthreads();
    // masked softmax over the legal cells (resnet.py:84-88: logits * mask + (1 - mask) * -1e9, softmax, * mask)
    float mx = -3.0e38f;
    for (int p = tid; p < 400; p += nthreads) {
        const int m = (p / 20) * BK_PAD_DIM + p % 20;
        const bool legal = (x64[size_t(m) * (BK_EVAL_IN_CH / 8)].z & 0xFFFFu) != 0u;
        if (logits) logits[p] = s_logit[p];
        if (!legal) s_logit[p] = -3.0e38f;
        mx = fmaxf(mx, s_logit[p]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(BK_FULL, mx, d));
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = s_red[0];
    for (int w = 1; w < nwarps; ++w) mx = fmaxf(mx, s_red[w]);
    __syncthreads();
    float sum = 0.0f;
    for (int p = tid; p < 400; p += nthreads) {
        const float l = s_logit[p];
        const float e = l > -1.0e38f ? expf(l - mx) : 0.0f;
        s_logit[p] = e;
        sum += e;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(BK_FULL, sum, d);
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    sum = 0.0f;
    for (int w = 0; w < nwarps; ++w) sum += s_red[w];
    const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;      // no legal cell: all zeros (the reference's softmax x mask)
    for (int p = tid; p < 400; p += nthreads) policy[p] = s_logit[p] * inv;
    // value head: Linear(400, 4) + tanh, softmax over the four seats (resnet.py:91-92)
    if (warp < 4) {
        float acc = 0.0f;
        for (int p = lane; p < 400; p += 32) acc += s_hv[p] * hp.lin_w[warp * 400 + p];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(BK_FULL, acc, d);
        if (lane == 0) s_red[8 + warp] = tanhf(acc + hp.lin_b[warp]);
    }
    __syncthreads();
    if (tid == 0) {
        const float t0 = s_red[8], t1 = s_red[9], t2 = s_red[10], t3 = s_red[11];
        const float tm = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
        const float e0 = expf(t0 - tm), e1 = expf(t1 - tm), e2 = expf(t2 - tm), e3 = expf(t3 - tm);
        const float es = e0 + e1 + e2 + e3;
        value[0] = e0 / es; value[1] = e1 / es; value[2] = e2 / es; value[3] = e3 / es;
        if (vtanh) { vtanh[0] = t0; vtanh[1] = t1; vtanh[2] = t2; vtanh[3] = t3; }
    }
}
