// Micro-benchmark (run under gpurun): how fast can a kernel write scattered 160-byte result rows straight into pinned
// host memory (zero-copy stores over PCIe), compared with cudaMemcpyAsync of the same rows as one contiguous range?
// Decides whether mrp_step_host can stream the rows of task-free envs out while the solver kernels still run.
// Outcome (B200 box, PCIe gen5): memcpy 53.9 GB/s, zero-copy 52.6 GB/s contiguous / 47 GB/s for a random 60 % of the rows;
// inside the step it lost all the same (measured dead ends).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/zc profiles/micro/zerocopy_rows.cu && /tmp/zc
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void k_out_rows(const float4* __restrict__ src, float4* __restrict__ dst, const int* __restrict__ list, int n, int q) {
    // q float4 per row; consecutive threads take consecutive float4 of consecutive listed rows
    const long total = (long)n * q;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int r = (int)(i / q), c = (int)(i % q);
        const long row = list ? list[r] : r;
        dst[row * q + c] = src[row * q + c];
    }
}

int main() {
    const int N = 1 << 20, q = 10;
    const size_t bytes = (size_t)N * q * sizeof(float4);
    float4 *dev, *host;
    cudaMalloc(&dev, bytes);
    cudaHostAlloc(&host, bytes, cudaHostAllocDefault);
    cudaMemset(dev, 1, bytes);
    std::vector<int> sub;
    srand(1);
    for (int i = 0; i < N; ++i) if (rand() % 100 < 60) sub.push_back(i);
    int* list;
    cudaMalloc(&list, sizeof(int) * sub.size());
    cudaMemcpy(list, sub.data(), sizeof(int) * sub.size(), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, size_t nbytes, auto&& fn) {
        for (int w = 0; w < 2; ++w) fn();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int it = 0; it < 10; ++it) fn();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        printf("%-44s %7.3f ms  %6.1f GB/s\n", name, ms, nbytes / ms * 1e-6);
    };
    time("cudaMemcpyAsync D2H, all rows", bytes, [&] { cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, 0); });
    for (int ctas : {16, 64, 148, 592}) for (int thr : {128, 512}) {
        char nm[96];
        snprintf(nm, sizeof nm, "zero-copy all rows, %d x %d", ctas, thr);
        time(nm, bytes, [&] { k_out_rows<<<ctas, thr>>>(dev, host, nullptr, N, q); });
        snprintf(nm, sizeof nm, "zero-copy 60%% listed rows, %d x %d", ctas, thr);
        time(nm, sub.size() * q * sizeof(float4), [&] { k_out_rows<<<ctas, thr>>>(dev, host, list, (int)sub.size(), q); });
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
