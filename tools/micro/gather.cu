// Microbenchmark: cost of per-lane random line gathers on B200 (cost model for the release-event structures).
// Each of 65536 threads loads `per` 16-byte words; variants: every word from a different random 128-byte line,
// or 4 words from one line.  Region size sweeps L2-resident .. multi-GB.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int LINES, int WORDS>
__global__ void gather(const uint4 *base, unsigned long long nlines, unsigned salt, uint4 *out) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int l = 0; l < LINES; l++) {
        const unsigned long long line = ((unsigned long long)hash32(tid * 16 + l + salt) * nlines) >> 32;
#pragma unroll
        for (int w = 0; w < WORDS; w++) {
            uint4 v = __ldcg(base + line * 8 + w);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    out[tid] = acc;
}
// per-env regions: thread t owns region t (stride bytes), picks random lines inside it (like the calendar)
template <int LINES, int WORDS>
__global__ void gather_own(const uint4 *base, unsigned lines_per_env, unsigned salt, uint4 *out) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int l = 0; l < LINES; l++) {
        const unsigned long long line = (unsigned long long)tid * lines_per_env + (hash32(tid * 16 + l + salt) % lines_per_env);
#pragma unroll
        for (int w = 0; w < WORDS; w++) {
            uint4 v = __ldcg(base + line * 8 + w);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    out[tid] = acc;
}
template <typename F>
float timeit(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(0); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < reps; r++) f(r + 1);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms * 1000.f / reps;
}
int main() {
    const int n = 65536, thr = 128, blocks = n / thr;
    uint4 *out; cudaMalloc(&out, n * sizeof(uint4));
    const size_t maxb = 8ULL << 30;
    uint4 *buf; if (cudaMalloc(&buf, maxb) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 1, maxb);
    printf("region_MB  1line_1w  1line_4w  2line_4w  4line_1w  4line_4w   (us per launch of 65536 threads)\n");
    for (size_t mb : {16, 64, 256, 512, 1024, 2048, 4096, 8192}) {
        unsigned long long nl = (mb << 20) / 128;
        float t0 = timeit([&](int r) { gather<1, 1><<<blocks, thr>>>(buf, nl, r * 7919u, out); }, 50);
        float t1 = timeit([&](int r) { gather<1, 4><<<blocks, thr>>>(buf, nl, r * 7919u, out); }, 50);
        float t2 = timeit([&](int r) { gather<2, 4><<<blocks, thr>>>(buf, nl, r * 7919u, out); }, 50);
        float t3 = timeit([&](int r) { gather<4, 1><<<blocks, thr>>>(buf, nl, r * 7919u, out); }, 50);
        float t4 = timeit([&](int r) { gather<4, 4><<<blocks, thr>>>(buf, nl, r * 7919u, out); }, 50);
        printf("%8zu  %8.2f  %8.2f  %8.2f  %8.2f  %8.2f\n", mb, t0, t1, t2, t3, t4);
    }
    printf("own-region (per-thread stride) lines_per_env: us for 1line_4w / 2line_4w / 4line_4w\n");
    for (unsigned lpe : {8, 64, 256, 512}) {
        float t1 = timeit([&](int r) { gather_own<1, 4><<<blocks, thr>>>(buf, lpe, r * 7919u, out); }, 50);
        float t2 = timeit([&](int r) { gather_own<2, 4><<<blocks, thr>>>(buf, lpe, r * 7919u, out); }, 50);
        float t4 = timeit([&](int r) { gather_own<4, 4><<<blocks, thr>>>(buf, lpe, r * 7919u, out); }, 50);
        printf("%8u (%5.0f MB)  %8.2f  %8.2f  %8.2f\n", lpe, lpe * 128.0 * n / 1048576.0, t1, t2, t4);
    }
    // empty kernel launch overhead reference
    float te = timeit([&](int r) { gather<1, 1><<<1, 32>>>(buf, 1, r, out); }, 50);
    printf("tiny launch: %.2f us\n", te);
    return 0;
}
