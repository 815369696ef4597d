// Dependent-chain latencies of the warp collectives and loads on the select's critical path (one warp, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 2048
__global__ void k(uint32_t* buf, const uint4* chase, long long* out, int iters) {
    const int lane = threadIdx.x;
    uint32_t v = buf[lane];
    long long t0, t1;
    // REDUX.MAX chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __reduce_max_sync(0xffffffffu, v + lane) ^ i;
    t1 = clock64(); if (lane == 0) out[0] = t1 - t0;
    // ballot + clz chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) v = (31u - __clz(__ballot_sync(0xffffffffu, (v + lane) & 1u) | 1u)) + v;
    t1 = clock64(); if (lane == 0) out[1] = t1 - t0;
    // shfl chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __shfl_sync(0xffffffffu, v, (v + i) & 31) + 1u;
    t1 = clock64(); if (lane == 0) out[2] = t1 - t0;
    // global pointer chase, 16-byte loads, all lanes same line (L1/L2 resident after first pass)
    uint32_t idx = v & 1023u;
    for (int i = 0; i < 64; ++i) idx = chase[idx].x;      // warm
    t0 = clock64();
    for (int i = 0; i < iters; ++i) idx = chase[idx + (lane & 1)].x;
    t1 = clock64(); if (lane == 0) out[3] = t1 - t0;
    // same through L2 only (ld.global.cg)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { uint4 r; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(chase + idx + (lane & 1))); idx = r.x; }
    t1 = clock64(); if (lane == 0) out[4] = t1 - t0;
    // shared-memory load chain
    __shared__ uint32_t sm[1024];
    for (int i = lane; i < 1024; i += 32) sm[i] = (i * 37 + 11) & 1023;
    __syncwarp();
    uint32_t s = idx & 1023u;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) s = sm[s];
    t1 = clock64(); if (lane == 0) out[5] = t1 - t0;
    // float chain: fmul + 2 fma + fmul + fadd
    float f = __uint_as_float(0x3f800000u | (v & 0xffff));
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { float q0 = f * 0.999f; f = __fmaf_rn(__fmaf_rn(-1.001f, q0, f), 0.999f, q0); f = f * 1.0001f + 0.5f; }
    t1 = clock64(); if (lane == 0) out[6] = t1 - t0;
    // store then load of the same global line by another lane (write-through / L1 behaviour)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { buf[32 + ((lane + 1) & 31)] = v + i; __syncwarp(); v = buf[32 + lane] + 1u; }
    t1 = clock64(); if (lane == 0) out[7] = t1 - t0;
    buf[lane] = v + idx + s + __float_as_uint(f);
}
int main() {
    uint32_t* buf; uint4* chase; long long* out;
    cudaMalloc(&buf, 4096); cudaMalloc(&chase, sizeof(uint4) * 1026); cudaMalloc(&out, 64);
    uint4 h[1026];
    for (int i = 0; i < 1026; ++i) { h[i].x = (i * 53 + 7) & 1023; h[i].y = h[i].z = h[i].w = 0; }
    cudaMemcpy(chase, h, sizeof h, cudaMemcpyHostToDevice); cudaMemset(buf, 1, 4096);
    const int iters = N;
    k<<<1, 32>>>(buf, chase, out, iters); cudaDeviceSynchronize();
    k<<<1, 32>>>(buf, chase, out, iters); cudaDeviceSynchronize();
    long long o[8]; cudaMemcpy(o, out, sizeof o, cudaMemcpyDeviceToHost);
    const char* names[8] = {"reduce_max_sync", "ballot+clz", "shfl_sync", "LDG.128 chase (L1)", "LDG.128 chase (.cg, L2)", "LDS chase", "fmul+2fma+fmul+fadd", "STG -> syncwarp -> LDG (other lane)"};
    for (int i = 0; i < 8; ++i) printf("%-40s %7.1f cycles per iteration\n", names[i], double(o[i]) / iters);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
