// bk_host.h — host-side plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/blokus_b200.h"
#include "bk_env_kernels.cuh"

void bk_set_error(const std::string& msg);
int bk_fail(int code, const std::string& msg);

#define BK_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return bk_fail(BK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
    } while (0)

// One spelling for kernel launches: <<<>>> under nvcc, the thread-per-lane harness under the
// tests' CPU warp emulator (tests/warp_emu, g++).
#ifdef BK_WARP_EMU
#define BK_LAUNCH(kernel, grid, block, stream, ...) emu_launch(int(grid), int(block), [&]() { kernel(__VA_ARGS__); })
#else
#define BK_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

struct bk_env {
    int n = 0;
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    BkState* d_states = nullptr;
    uint16_t* d_hist = nullptr;      // [n][BK_HIST_CAP]: tile | player << 9
    int32_t* d_i32 = nullptr;        // scratch [n][4]: tiles / finish / status / steps
    uint64_t* d_hash = nullptr;      // [n]
    unsigned long long* d_counters = nullptr;  // [8]
    BkSummary* d_summary = nullptr;  // [n]
    uint8_t* d_bytes = nullptr;      // [n][2000] staging for masks / boards / planes
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t uev[2] = {nullptr, nullptr};  // caller-driven region timing
    float last_ms = 0.0f;
    bool timing_pending = false;
    bool borrowed = false;           // owned by a bk_selfplay
};

// implemented in bk_env.cu, used by bk_mcts.cu
int bk_env_alloc(int n_games, int device, cudaStream_t stream, bk_env** out);

// implemented in bk_conv.cu, used by bk_eval.cu: one 3x3 convolution on the padded NHWC bf16 layout
int bk_conv_launch(const void* dev_x, const void* dev_w, const float* dev_bias, const void* dev_residual, void* dev_y,
                   int batch, int in_channels, int relu, void* cuda_stream);

// The policy/value network of model/resnet.py:44-94 (eval mode, BatchNorm folded) resident on one device:
// weights, the three ping-pong activation matrices and the first layer's input, all in the convolution
// kernel's padded NHWC bf16 layout.  Defined here because bk_mcts.cu drives it inside the self-play loop.
struct bk_evaluator {
    int device = 0;
    int blocks = 0;
    int cap_rows = 0;                 // positions per forward pass the buffers hold
    uint16_t* d_w_in = nullptr;       // bf16 [9][256][64]
    float* d_b_in = nullptr;          // [256]
    uint16_t* d_w_blk = nullptr;      // bf16 [2*blocks][9][256][256]
    float* d_b_blk = nullptr;         // [2*blocks][256]
    float* d_head = nullptr;          // head_w [2][256], head_affine [4], lin_w [4][400], lin_b [4]
    uint16_t* d_act[3] = {nullptr, nullptr, nullptr};   // bf16 [cap_rows*441][256]
    uint16_t* d_x64 = nullptr;        // bf16 [cap_rows*441][64]
    float* d_policy = nullptr;        // [cap_rows][400]   (outputs of the fused self-play loop)
    float* d_value = nullptr;         // [cap_rows][4]
};

// implemented in bk_eval.cu: the network on the positions already written to ev->d_x64 (rows of them)
int bk_evaluator_forward_x64(bk_evaluator* ev, int rows, float* dev_policy, float* dev_value, float* dev_logits,
                             float* dev_vtanh, cudaStream_t stream);
