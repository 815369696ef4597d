// bk_eval.cu — the leaf evaluator of BASELINE.json config 4 as ONE native object: the reference's policy/value
// ResNet (model/resnet.py:44-94) in eval mode, every layer on this library's kernels — input packing, the
// 2*blocks + 1 tcgen05 convolutions (bk_conv.cu) and the fused heads (bk_eval_kernels.cuh).  No PyTorch between
// the tree kernels' planes and the policy/value rows bk_selfplay_expand_backup consumes.
#include <vector>

#include "bk_host.h"
#include "bk_eval_kernels.cuh"

__global__ void __launch_bounds__(256) k_eval_pack_planes(const float* __restrict__ planes, uint16_t* __restrict__ x64, int rows) {
    const int row = blockIdx.x;
    if (row >= rows) return;
    kb_pack_planes(planes + size_t(row) * 2000, reinterpret_cast<uint4*>(x64 + size_t(row) * BK_PAD_IMAGE * BK_EVAL_IN_CH),
                   threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(256) k_eval_heads(const uint16_t* __restrict__ act, const uint16_t* __restrict__ x64, BkHeadParams hp,
                                                    int rows, float* __restrict__ policy, float* __restrict__ value,
                                                    float* __restrict__ logits, float* __restrict__ vtanh) {
    __shared__ float sh[816];
    const int row = blockIdx.x;
    if (row >= rows) return;
    kb_heads(reinterpret_cast<const uint4*>(act + size_t(row) * BK_PAD_IMAGE * BK_EVAL_CH),
             reinterpret_cast<const uint4*>(x64 + size_t(row) * BK_PAD_IMAGE * BK_EVAL_IN_CH), hp, policy + size_t(row) * 400,
             value + size_t(row) * 4, logits ? logits + size_t(row) * 400 : nullptr, vtanh ? vtanh + size_t(row) * 4 : nullptr, sh,
             threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(256) k_env_planes_nhwc(const BkState* __restrict__ states, uint16_t* __restrict__ x64, int n) {
    const int g = blockIdx.x;
    if (g >= n) return;
    kb_planes_nhwc(&states[g], reinterpret_cast<uint4*>(x64 + size_t(g) * BK_PAD_IMAGE * BK_EVAL_IN_CH), threadIdx.x, blockDim.x);
}

static BkHeadParams head_params(const float* d_head) {
    BkHeadParams hp;
    hp.head_w = d_head;
    hp.head_affine = d_head + 512;
    hp.lin_w = d_head + 516;
    hp.lin_b = d_head + 516 + 1600;
    return hp;
}

int bk_evaluator_forward_x64(bk_evaluator* ev, int rows, float* dev_policy, float* dev_value, float* dev_logits,
                             float* dev_vtanh, cudaStream_t st) {
    if (rows <= 0) return BK_OK;
    const size_t wsz = size_t(9) * BK_EVAL_CH * BK_EVAL_CH;
    uint16_t *a = ev->d_act[0], *t = ev->d_act[1], *b = ev->d_act[2];
    int rc = bk_conv_launch(ev->d_x64, ev->d_w_in, ev->d_b_in, nullptr, a, rows, BK_EVAL_IN_CH, 0, st);   // model.input: no BN / ReLU
    for (int i = 0; i < ev->blocks && !rc; ++i) {
        rc = bk_conv_launch(a, ev->d_w_blk + size_t(2 * i) * wsz, ev->d_b_blk + size_t(2 * i) * BK_EVAL_CH, nullptr, t, rows, BK_EVAL_CH, 1, st);
        if (!rc) rc = bk_conv_launch(t, ev->d_w_blk + size_t(2 * i + 1) * wsz, ev->d_b_blk + size_t(2 * i + 1) * BK_EVAL_CH, a, b, rows, BK_EVAL_CH, 1, st);
        uint16_t* s = a; a = b; b = s;
    }
    if (rc) return rc;
    BK_LAUNCH(k_eval_heads, rows, 256, st, a, ev->d_x64, head_params(ev->d_head), rows, dev_policy, dev_value, dev_logits, dev_vtanh);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

extern "C" {

int bk_env_board_state_nhwc(bk_env* env, void* dev_x64) {
    if (!env || !dev_x64) return bk_fail(BK_ERR_INVALID_ARG, "bk_env_board_state_nhwc: null argument");
    BK_CUDA(cudaSetDevice(env->device));
    BK_LAUNCH(k_env_planes_nhwc, env->n, 256, env->stream, env->d_states, static_cast<uint16_t*>(dev_x64), env->n);
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaStreamSynchronize(env->stream));
    return BK_OK;
}

int bk_eval_pack_planes(const float* dev_planes, int rows, void* dev_x64, void* cuda_stream) {
    if (!dev_planes || !dev_x64 || rows <= 0) return bk_fail(BK_ERR_INVALID_ARG, "bk_eval_pack_planes: bad argument");
    BK_LAUNCH(k_eval_pack_planes, rows, 256, static_cast<cudaStream_t>(cuda_stream), dev_planes, static_cast<uint16_t*>(dev_x64), rows);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

int bk_eval_heads(const void* dev_act, const void* dev_x64, const float* dev_head_params, int rows, float* dev_policy,
                  float* dev_value, float* dev_logits, float* dev_vtanh, void* cuda_stream) {
    if (!dev_act || !dev_x64 || !dev_head_params || !dev_policy || !dev_value || rows <= 0)
        return bk_fail(BK_ERR_INVALID_ARG, "bk_eval_heads: bad argument");
    BK_LAUNCH(k_eval_heads, rows, 256, static_cast<cudaStream_t>(cuda_stream), static_cast<const uint16_t*>(dev_act),
              static_cast<const uint16_t*>(dev_x64), head_params(dev_head_params), rows, dev_policy, dev_value, dev_logits, dev_vtanh);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

void bk_evaluator_destroy(bk_evaluator* ev) {
    if (!ev) return;
    cudaSetDevice(ev->device);
    cudaFree(ev->d_w_in); cudaFree(ev->d_b_in); cudaFree(ev->d_w_blk); cudaFree(ev->d_b_blk); cudaFree(ev->d_head);
    for (uint16_t* p : ev->d_act) cudaFree(p);
    cudaFree(ev->d_x64); cudaFree(ev->d_policy); cudaFree(ev->d_value);
    delete ev;
}

int bk_evaluator_create(int device, int blocks, int max_rows, const void* w_in, const float* b_in, const void* w_blocks,
                        const float* b_blocks, const float* head_w, const float* head_affine, const float* lin_w,
                        const float* lin_b, bk_evaluator** out) {
    if (!out || blocks < 0 || blocks > 256 || max_rows <= 0 || !w_in || !b_in || (blocks && (!w_blocks || !b_blocks)) || !head_w ||
        !head_affine || !lin_w || !lin_b)
        return bk_fail(BK_ERR_INVALID_ARG, "bk_evaluator_create: bad argument");
    BK_CUDA(cudaSetDevice(device));
    bk_evaluator* ev = new bk_evaluator();
    ev->device = device; ev->blocks = blocks; ev->cap_rows = max_rows;
    const size_t wsz = size_t(9) * BK_EVAL_CH * BK_EVAL_CH, m = size_t(max_rows) * BK_PAD_IMAGE;
    auto fail = [&](cudaError_t e, const char* what) {
        bk_evaluator_destroy(ev);
        return bk_fail(BK_ERR_CUDA, std::string("bk_evaluator_create: ") + what + ": " + cudaGetErrorString(e));
    };
    cudaError_t e;
#define BK_EV_ALLOC(ptr, bytes) if ((e = cudaMalloc(&(ptr), (bytes))) != cudaSuccess) return fail(e, #ptr)
    BK_EV_ALLOC(ev->d_w_in, size_t(9) * BK_EVAL_CH * BK_EVAL_IN_CH * 2);
    BK_EV_ALLOC(ev->d_b_in, BK_EVAL_CH * sizeof(float));
    BK_EV_ALLOC(ev->d_w_blk, (blocks ? size_t(2 * blocks) * wsz : 1) * 2);
    BK_EV_ALLOC(ev->d_b_blk, (blocks ? size_t(2 * blocks) * BK_EVAL_CH : 1) * sizeof(float));
    BK_EV_ALLOC(ev->d_head, size_t(516 + 1604) * sizeof(float));
    for (int i = 0; i < 3; ++i) BK_EV_ALLOC(ev->d_act[i], m * BK_EVAL_CH * 2);
    BK_EV_ALLOC(ev->d_x64, m * BK_EVAL_IN_CH * 2);
    BK_EV_ALLOC(ev->d_policy, size_t(max_rows) * 400 * sizeof(float));
    BK_EV_ALLOC(ev->d_value, size_t(max_rows) * 4 * sizeof(float));
#undef BK_EV_ALLOC
    // channels 5..63 and the pad row / column of every image of the input are zeros for good: the writers
    // (kb_planes_nhwc, kb_pack_planes) only ever touch channels 0..7 of the 400 real cells
    if ((e = cudaMemset(ev->d_x64, 0, m * BK_EVAL_IN_CH * 2)) != cudaSuccess) return fail(e, "memset");
    std::vector<float> head(516 + 1604);
    for (int i = 0; i < 512; ++i) head[size_t(i)] = head_w[i];
    for (int i = 0; i < 4; ++i) head[512 + size_t(i)] = head_affine[i];
    for (int i = 0; i < 1600; ++i) head[516 + size_t(i)] = lin_w[i];
    for (int i = 0; i < 4; ++i) head[2116 + size_t(i)] = lin_b[i];
    if ((e = cudaMemcpy(ev->d_w_in, w_in, size_t(9) * BK_EVAL_CH * BK_EVAL_IN_CH * 2, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "copy");
    if ((e = cudaMemcpy(ev->d_b_in, b_in, BK_EVAL_CH * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "copy");
    if (blocks) {
        if ((e = cudaMemcpy(ev->d_w_blk, w_blocks, size_t(2 * blocks) * wsz * 2, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "copy");
        if ((e = cudaMemcpy(ev->d_b_blk, b_blocks, size_t(2 * blocks) * BK_EVAL_CH * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "copy");
    }
    if ((e = cudaMemcpy(ev->d_head, head.data(), head.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "copy");
    *out = ev;
    return BK_OK;
}

int bk_evaluator_max_rows(const bk_evaluator* ev) { return ev ? ev->cap_rows : 0; }

int bk_evaluator_forward(bk_evaluator* ev, const float* dev_planes, int rows, float* dev_policy, float* dev_value,
                         float* dev_logits, float* dev_vtanh, void* cuda_stream) {
    if (!ev || !dev_planes || !dev_policy || !dev_value) return bk_fail(BK_ERR_INVALID_ARG, "bk_evaluator_forward: null argument");
    if (rows <= 0 || rows > ev->cap_rows) return bk_fail(BK_ERR_INVALID_ARG, "bk_evaluator_forward: rows must be in 1..max_rows");
    BK_CUDA(cudaSetDevice(ev->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    BK_LAUNCH(k_eval_pack_planes, rows, 256, st, dev_planes, ev->d_x64, rows);
    BK_CUDA(cudaGetLastError());
    return bk_evaluator_forward_x64(ev, rows, dev_policy, dev_value, dev_logits, dev_vtanh, st);
}

}  // extern "C"
