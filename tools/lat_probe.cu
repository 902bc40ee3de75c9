// Dependent-issue latency of the instructions on the demodulator's per-symbol feedback chain, one warp on one SM:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/lat_probe.bin tools/lat_probe.cu
// Prints cycles per dependent operation.  The figures decide which arithmetic the Costas phase chain / sin-cos may use.
#include <cstdio>
#include <cuda_runtime.h>

#define CHAIN 256
#define PROBE(name, type, init, body)                                                                   \
    __global__ void k_##name(type *out, long long *cyc, type a, type b)                                 \
    {                                                                                                   \
        type v = init;                                                                                  \
        long long t0 = clock64();                                                                       \
        _Pragma("unroll") for (int i = 0; i < CHAIN; i++) { body; }                                     \
        long long t1 = clock64();                                                                       \
        out[threadIdx.x] = v;                                                                           \
        if (threadIdx.x == 0) *cyc = t1 - t0;                                                           \
    }

PROBE(dadd, double, a, v = __dadd_rn(v, b))
PROBE(dmul, double, a, v = __dmul_rn(v, b))
PROBE(dfma, double, a, v = __fma_rn(v, b, a))
PROBE(fadd, float, a, v = __fadd_rn(v, b))
PROBE(fmul, float, a, v = __fmul_rn(v, b))
PROBE(ffma, float, a, v = __fmaf_rn(v, b, a))
PROBE(iadd, int, a, v = (v + b) ^ a)
PROBE(d2f2d, double, a, v = (double)(float)v + b)
PROBE(frsq, float, a, v = rsqrtf(v) + b)
PROBE(shfl, float, a, v = __shfl_xor_sync(0xffffffffu, v, 1) + b)
PROBE(dsetp, double, a, if (v > b) v = __dadd_rn(v, -b); else v = __dadd_rn(v, a))

__global__ void k_lds(int *out, long long *cyc, int a, int b)
{
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) s[i] = (i + 33) & 1023;
    __syncwarp();
    int v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < CHAIN; i++) v = s[v];
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

// throughput: every SM full of warps, 8 independent chains per thread
template <typename T>
__global__ void k_tput(T *out, T a, T b, int iters)
{
    T v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = a + (T)(threadIdx.x + j);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = v[j] * b + a;   // -fmad=false: a multiply and an add
    }
    T s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T>
void tput(const char *name)
{
    T *out;
    cudaMalloc(&out, 148 * 2 * 1024 * sizeof(T));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096;
    k_tput<T><<<148 * 2, 1024>>>(out, (T)1.0, (T)0.999, iters);
    cudaEventRecord(e0);
    k_tput<T><<<148 * 2, 1024>>>(out, (T)1.0, (T)0.999, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 148.0 * 2 * 1024 * 8 * 2 * iters;   // thread-level mul + add
    printf("%-8s %.2f T thread-ops/s = %.1f warp-instructions per SM per cycle at 1.965 GHz\n", name, ops / ms / 1e9,
           ops / 32 / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}

template <typename T, typename K>
void run(const char *name, K kernel, T a, T b)
{
    T *out;
    long long *cyc, h = 0;
    cudaMalloc(&out, 32 * sizeof(T));
    cudaMalloc(&cyc, sizeof(long long));
    for (int r = 0; r < 3; r++) kernel<<<1, 32>>>(out, cyc, a, b);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-8s %.1f cycles per dependent operation\n", name, (double)h / CHAIN);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    run<double>("DADD", k_dadd, 1.0, 1e-3);
    run<double>("DMUL", k_dmul, 1.0, 1.0000001);
    run<double>("DFMA", k_dfma, 1.0, 0.999);
    run<float>("FADD", k_fadd, 1.0f, 1e-3f);
    run<float>("FMUL", k_fmul, 1.0f, 1.0001f);
    run<float>("FFMA", k_ffma, 1.0f, 0.999f);
    run<int>("IADD+XOR", k_iadd, 1, 3);
    run<double>("D2F+F2D+DADD", k_d2f2d, 1.0, 1e-3);
    run<float>("RSQ+FADD", k_frsq, 1.0f, 1e-3f);
    run<float>("SHFL+FADD", k_shfl, 1.0f, 1e-3f);
    run<double>("DSETP+DADD", k_dsetp, 1.0, 0.5);
    run<int>("LDS", k_lds, 0, 0);
    tput<double>("FP64 mul+add");
    tput<float>("FP32 mul+add");
    return cudaDeviceSynchronize() != cudaSuccess;
}
