// Micro-benchmark behind DESIGN.md K7: latency of dependent fp64 operations in ONE warp on this GPU (cycles per operation).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/fp64_latency tools/fp64_latency.cu && tools/_bin/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double seed, long long* out, double* sink) {
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    double x = seed + lane * 1e-3;
    long long t0, t1;
    constexpr int N = 256;
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = fma(x, 1.0000001, 1e-9);
    t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = sqrt(x + 2.0);
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = 3.0 / (x + 1.0);
    t1 = clock64();
    if (lane == 0) out[2] = (t1 - t0) / N;
    float y = (float)x;
    t0 = clock64();
    for (int i = 0; i < N; ++i) y = sqrtf(y + 2.0f);
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) y = __fdiv_rn(3.0f, y + 1.0f);
    t1 = clock64();
    if (lane == 0) out[4] = (t1 - t0) / N;
    x += y;
    t0 = clock64();
    for (int i = 0; i < N; ++i) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31] + 1.0; __syncwarp(); }
    t1 = clock64();
    if (lane == 0) out[5] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = __shfl_down_sync(0xffffffffu, x, 1) + 1.0;
    t1 = clock64();
    if (lane == 0) out[6] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = x * 1.0000001;
    t1 = clock64();
    if (lane == 0) out[7] = (t1 - t0) / N;
    sink[lane] = x;
}
int main() {
    long long* out; double* sink;
    cudaMallocManaged(&out, 8 * sizeof(long long)); cudaMalloc(&sink, 64 * sizeof(double));
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(1.5, out, sink); cudaDeviceSynchronize(); }
    const char* names[8] = {"dfma", "dsqrt", "ddiv", "fsqrt", "fdiv_rn", "smem+2 syncwarp+dadd", "shfl(double)+dadd", "dmul"};
    for (int i = 0; i < 8; ++i) printf("%-24s %lld cycles per dependent op\n", names[i], out[i]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
