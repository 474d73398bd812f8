// Probe (development aid): cost of a data-with-flag ("LL") all-to-all hand-over inside one persistent kernel.
// Round r: every CTA writes its slice of y as 8-byte (half2, epoch) packets; every CTA then reads ALL of y,
// each thread spinning on its own packets.  No fences, no counters, no barriers.  us per round.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldv4(const void *p) {
    uint4 r;
#ifdef GPU_SCOPE
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
#else
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
#endif
    return r;
}
__device__ __forceinline__ void stv2(void *p, uint32_t a, uint32_t b) {
#ifdef GPU_SCOPE
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
#else
    asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
#endif
}

// n = number of halfs in y (n/2 packets).  poll_mode 0: every thread spins on its vectors; 1: lane 0 spins first
__global__ void __launch_bounds__(544, 1) ll_rounds(uint2 *buf, uint32_t n, int rounds, int poll_mode, uint32_t work, float *sink,
                                                    unsigned long long *cycles, uint32_t *dbg) {
    const uint32_t G = gridDim.x, npk = n / 2;
    const uint32_t per = (npk + G - 1) / G;
    uint32_t acc = 0;
    float f = 1.f;
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) t0 = clock64();
    for (int r = 1; r <= rounds; r++) {
        uint2 *b = buf + (size_t)(r & 1) * npk;
        // "compute" (dependent on what we read last round)
        for (uint32_t i = 0; i < work; i++) f = f * 1.0001f + (float)(acc & 1);
        const uint32_t p0 = blockIdx.x * per;
        for (uint32_t i = threadIdx.x; i < per && p0 + i < npk; i += blockDim.x) stv2(b + p0 + i, (uint32_t)r + (f == 7.f), (uint32_t)r);
        // read all: thread t reads vectors t, t+512, ... (2 packets each)
        if (threadIdx.x < 512) {
            if (poll_mode == 1) {
                if ((threadIdx.x & 31) == 0) {
                    long long tw = clock64();
                    while (ldv4(b + 2 * threadIdx.x).y != (uint32_t)r) {
                        __nanosleep(40);
                        if (clock64() - tw > 2000000000ll) { if (atomicAdd(dbg, 1u) == 0u) { dbg[1] = r; dbg[2] = 0xffffffffu; dbg[5] = blockIdx.x; dbg[6] = threadIdx.x; } __trap(); }
                    }
                }
                __syncwarp();
            }
            for (uint32_t v = threadIdx.x; v < npk / 2; v += 512) {
                uint4 q;
                long long tw = 0;
                while (true) {
                    q = ldv4(b + 2 * v);
                    if (q.y == (uint32_t)r && q.w == (uint32_t)r) break;
                    __nanosleep(20);
                    if (tw == 0) tw = clock64();
                    else if (clock64() - tw > 2000000000ll) {  // ~1 s: report and bail out of the whole kernel
                        if (atomicAdd(dbg, 1u) == 0u) { dbg[1] = r; dbg[2] = v; dbg[3] = q.y; dbg[4] = q.w; dbg[5] = blockIdx.x; dbg[6] = threadIdx.x; }
                        __trap();
                    }
                }
                acc += q.x + q.z;
            }
        }
        __syncthreads();  // a job's epilogue starts after all warps of the CTA are done with the job (and its reads)
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
    if (threadIdx.x == 0) sink[blockIdx.x] = (float)acc + f;
}

int main() {
    int sms = 0;
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint2 *buf;
    float *sink;
    unsigned long long *cycles;
    const uint32_t nmax = 32768;
    CK(cudaMalloc(&buf, 2 * (nmax / 2) * 8));
    CK(cudaMalloc(&sink, 1024 * 4));
    CK(cudaMalloc(&cycles, 8));
    uint32_t *dbg;
    CK(cudaMalloc(&dbg, 64));
    CK(cudaMemset(dbg, 0, 64));
    setvbuf(stdout, NULL, _IONBF, 0);
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int rounds = 500;
    for (uint32_t n : {4096u, 14336u, 28672u}) {
        for (int pm = 0; pm < 2; pm++) {
            for (uint32_t work : {0u, 2000u}) {
                CK(cudaMemsetAsync(buf, 0, 2 * (nmax / 2) * 8, s));
                void *args[] = {&buf, &n, &rounds, &pm, &work, &sink, &cycles, &dbg};
                CK(cudaEventRecord(e0, s));
                CK(cudaLaunchCooperativeKernel((void *)ll_rounds, dim3(sms), dim3(544), args, 0, s));
                CK(cudaEventRecord(e1, s));
                cudaError_t se = cudaStreamSynchronize(s);
                if (se != cudaSuccess) {
                    printf("kernel failed: %s (n=%u pm=%d work=%u)\n", cudaGetErrorString(se), n, pm, work);
                    return 1;
                }
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                unsigned long long cyc = 0;
                CK(cudaMemcpy(&cyc, cycles, 8, cudaMemcpyDeviceToHost));
                printf("LL hand-over n=%u halfs poll_mode=%d work=%u: %.3f us/round (%.0f cycles)\n", n, pm, work, ms * 1e3 / rounds,
                       (double)cyc / rounds);
            }
        }
    }
    return 0;
}
