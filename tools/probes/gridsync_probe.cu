// Probe (development aid, not product): what does one grid-wide hand-over cost on B200?
//   A. persistent kernel, 148 CTAs x 544 threads: per round every CTA writes a slice of y, releases a counter,
//      acquires it when all CTAs arrived, then reads ALL of y (ld.cg) — the x hand-over of a persistent GEMV chain
//   B. the same rounds as a CUDA graph of PDL-chained kernels (what round 1 does)
// prints us per round for both.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(uint32_t *p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_cg(const float *p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// mode 0: counter only.  mode 1: + y exchange (n floats total, every CTA reads all of them)
__global__ void __launch_bounds__(544, 1) persistent_rounds(uint32_t *counter, float *y, uint32_t n, int rounds, int mode,
                                                          float *sink, unsigned long long *cycles) {
    const uint32_t G = gridDim.x;
    float acc = 0.f;
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) t0 = clock64();
    for (int r = 1; r <= rounds; r++) {
        float *yr = y + (size_t)(r & 1) * n;  // double buffer: round r writes buffer r&1, reads it after the barrier
        if (mode == 1) {
            for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += G * blockDim.x) yr[i] = (float)r + acc * 0.f;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            red_release(counter, 1u);  // release: the CTA's y stores (ordered by bar.sync + cumulativity) become visible
            while (ld_acquire(counter) < (uint32_t)r * G) {}
        }
        __syncthreads();
        if (mode == 1) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) acc += ld_cg(yr + i);
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
    if (acc == 12345.678f) *sink = acc;
    if (mode == 1 && threadIdx.x == 0) sink[1 + blockIdx.x] = acc;
}

__global__ void __launch_bounds__(544, 1) pdl_round(float *y, uint32_t n, int r, float *sink) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    float acc = 0.f;
    const float *yp = y + (size_t)((r - 1) & 1) * n;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) acc += __ldcg(yp + i);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    float *yr = y + (size_t)(r & 1) * n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) yr[i] = (float)r + acc * 0.f;
    if (acc == 12345.678f) *sink = acc;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const uint32_t n = 4096;
    uint32_t *counter;
    float *y, *sink;
    unsigned long long *cycles;
    CK(cudaMalloc(&counter, 4));
    CK(cudaMalloc(&y, 2 * n * 4));
    CK(cudaMalloc(&sink, (2 + 1024) * 4));
    CK(cudaMalloc(&cycles, 8));
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int rounds = 2000;
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemsetAsync(counter, 0, 4, s));
            CK(cudaMemsetAsync(y, 0, 2 * n * 4, s));
            void *args[] = {&counter, &y, (void *)&n, (void *)&rounds, &mode, &sink, &cycles};
            CK(cudaEventRecord(e0, s));
            CK(cudaLaunchCooperativeKernel((void *)persistent_rounds, dim3(sms), dim3(544), args, 0, s));
            CK(cudaEventRecord(e1, s));
            CK(cudaStreamSynchronize(s));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            unsigned long long cyc = 0;
            CK(cudaMemcpy(&cyc, cycles, 8, cudaMemcpyDeviceToHost));
            float chk[2];
            CK(cudaMemcpy(chk, sink + 1, 8, cudaMemcpyDeviceToHost));
            printf("persistent mode %d (%s): %.3f us/round (%.0f cycles/round), check %.1f (expect %.1f)\n", mode,
                   mode ? "counter + 16 KB y exchange" : "counter only", ms * 1e3 / rounds, (double)cyc / rounds, chk[0],
                   mode ? (double)n * rounds * (rounds + 1) / 2.0 : 0.0);
        }
    }
    // PDL chain in a graph
    {
        const int R = 256;
        cudaGraph_t g;
        cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal));
        for (int r = 1; r <= R; r++) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(sms);
            cfg.blockDim = dim3(544);
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, pdl_round, y, n, r, sink));
        }
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int i = 0; i < 3; i++) CK(cudaGraphLaunch(ge, s));
        CK(cudaEventRecord(e0, s));
        for (int i = 0; i < 10; i++) CK(cudaGraphLaunch(ge, s));
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("PDL graph chain (148 CTAs x 544 thr, 16 KB y hand-over): %.3f us/kernel\n", ms * 1e3 / (10.0 * R));
    }
    return 0;
}
