// PCIe pipeline probe: H2D of 8 fields and D2H of 6 fields in row-block chunks, concurrently, as aerobulk_gpu_model does.
// A: one cudaMemcpyAsync per field and chunk.  B: one cudaMemcpy2DAsync per chunk (fields equally spaced in one slab).
// build: nvcc -O2 -o duplex_probe duplex_probe.cu ; run: ./duplex_probe [n] [chunks]
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__global__ void busy(double *p, long long n, int iters)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = p[i];
    for (int k = 0; k < iters; ++k) v = fma(v, 1.0000001, 1e-9);
    p[i] = v;
}
int main(int argc, char **argv)
{
    const long long n = argc > 1 ? atoll(argv[1]) : 1036800;
    const int K = argc > 2 ? atoi(argv[2]) : 6;
    double *hin, *hout, *din, *dout;
    CK(cudaHostAlloc(&hin, 8 * n * 8, cudaHostAllocDefault));
    CK(cudaHostAlloc(&hout, 6 * n * 8, cudaHostAllocDefault));
    CK(cudaMalloc(&din, 8 * n * 8));
    CK(cudaMalloc(&dout, 6 * n * 8));
    for (long long i = 0; i < 8 * n; ++i) hin[i] = 1.0;
    cudaStream_t si, sk, so;
    CK(cudaStreamCreateWithFlags(&si, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> ei(K), ek(K);
    for (int c = 0; c < K; ++c) { CK(cudaEventCreateWithFlags(&ei[c], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ek[c], cudaEventDisableTiming)); }
    std::vector<long long> cs(K + 1, 0);
    {
        double wsum = K * (K + 1) / 2.0, acc = 0;
        for (int c = 0; c < K; ++c) { acc += K - c; cs[c + 1] = (long long)(n * acc / wsum) / 2048 * 2048; }
        cs[K] = n;
    }
    for (int mode = 0; mode < 2; ++mode) {
        double best = 1e9;
        for (int rep = 0; rep < 12; ++rep) {
            CK(cudaDeviceSynchronize());
            auto t0 = std::chrono::steady_clock::now();
            for (int c = 0; c < K; ++c) {
                const long long s0 = cs[c], len = cs[c + 1] - s0;
                if (mode == 0)
                    for (int k = 0; k < 8; ++k) CK(cudaMemcpyAsync(din + k * n + s0, hin + k * n + s0, len * 8, cudaMemcpyHostToDevice, si));
                else
                    CK(cudaMemcpy2DAsync(din + s0, n * 8, hin + s0, n * 8, len * 8, 8, cudaMemcpyHostToDevice, si));
                CK(cudaEventRecord(ei[c], si));
            }
            for (int c = 0; c < K; ++c) {
                const long long s0 = cs[c], len = cs[c + 1] - s0;
                CK(cudaStreamWaitEvent(sk, ei[c], 0));
                busy<<<(unsigned)((len + 255) / 256), 256, 0, sk>>>(dout + s0, len, 6000);   // ~0.7 ms per 1M points
                CK(cudaEventRecord(ek[c], sk));
                CK(cudaStreamWaitEvent(so, ek[c], 0));
                if (mode == 0)
                    for (int k = 0; k < 6; ++k) CK(cudaMemcpyAsync(hout + k * n + s0, dout + k * n + s0, len * 8, cudaMemcpyDeviceToHost, so));
                else
                    CK(cudaMemcpy2DAsync(hout + s0, n * 8, dout + s0, n * 8, len * 8, 6, cudaMemcpyDeviceToHost, so));
            }
            CK(cudaStreamSynchronize(sk));
            CK(cudaStreamSynchronize(so));
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (rep > 1 && ms < best) best = ms;
        }
        printf("%s: n=%lld chunks=%d  %.3f ms per call (%.1f GB/s aggregate)\n", mode ? "2-D copies (1 per chunk) " : "1-D copies (1 per field)", n, K, best, 14.0 * n * 8 / best / 1e6);
    }
    return 0;
}
