// TEST HARNESS ONLY — see cuda_shim.h.
#include <chrono>

#include "cuda_shim.h"

thread_local EmuIdx threadIdx, blockIdx, blockDim, gridDim;
thread_local EmuWarp* emu_warp = nullptr;
thread_local EmuCta* emu_cta = nullptr;
thread_local int emu_lane = 0;
thread_local unsigned emu_phase = 0;

void emu_launch(int grid, int block, const std::function<void()>& body) {
    if (block % 32 != 0 && block > 32) std::abort();
    for (int b = 0; b < grid; ++b) {
        EmuCta cta;
        cta.bar = std::make_unique<std::barrier<>>(block);
        const int nwarps = (block + 31) / 32;
        for (int w = 0; w < nwarps; ++w) cta.warps.emplace_back(std::make_unique<EmuWarp>());
        std::vector<std::thread> th;
        th.reserve(size_t(block));
        for (int t = 0; t < block; ++t) {
            th.emplace_back([&, t]() {
                threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
                emu_cta = &cta;
                emu_warp = cta.warps[size_t(t / 32)].get();
                emu_lane = t % 32;
                emu_phase = 0;
                body();
                // a thread that returns early must keep later collectives/barriers from deadlocking
                emu_warp->bar.arrive_and_drop();
                cta.bar->arrive_and_drop();
            });
        }
        for (auto& t : th) t.join();
    }
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent{0.0}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = now_ms(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = float(b->t - a->t); return cudaSuccess; }
