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
