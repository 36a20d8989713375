// Micro-benchmark: issue rate of DMMA.8x8x4 on sm_100a, alone and mixed with the DMUL / LDS traffic the
// Hessian kernel needs.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench.bin dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int V>
__global__ void bench(double* out, const double* in, int iters) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double acc[4][4][2];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double af[4], bf[4];
  for (int i = 0; i < 4; ++i) { af[i] = in[lane + 32 * i]; bf[i] = in[lane + 32 * i + 128]; }
  double wk = in[300];
  for (int it = 0; it < iters; ++it) {
    if (V & 2) {
      const double2* p = reinterpret_cast<const double2*>(sm) + ((it & 7) * 128 + lane);
      double2 a0 = p[0], a1 = p[32], b0 = p[64], b1 = p[96];
      af[0] = a0.x; af[1] = a0.y; af[2] = a1.x; af[3] = a1.y;
      bf[0] = b0.x; bf[1] = b0.y; bf[2] = b1.x; bf[3] = b1.y;
      wk = sm[2048 + ((it * 4 + (lane & 3)) & 1023)];
    }
    double a2[4] = {af[0], af[1], af[2], af[3]};
    if (V & 1) {
      for (int i = 0; i < 4; ++i) a2[i] = af[i] * wk;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a2[i], bf[j]);
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(int threads, int ctas_per_sm, const char* name, double* out, double* in) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<V><<<sms * ctas_per_sm, threads>>>(out, in, 100);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    bench<V><<<sms * ctas_per_sm, threads>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double warps = (double)sms * ctas_per_sm * threads / 32;
  const double flops = warps * iters * 16.0 * 512.0;
  const double clk = 1.965e9;
  const double dmma_per_smsp = (double)ctas_per_sm * threads / 32 / 4 * iters * 16.0;
  printf("%-28s threads=%4d ctas/sm=%d  %.3f ms  %.2f TFLOP/s  %.2f clk/DMMA/SMSP (at 1965 MHz)\n", name, threads, ctas_per_sm,
         best, flops / (best * 1e-3) / 1e12, best * 1e-3 * clk / dmma_per_smsp);
}

int main() {
  double *out, *in;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  cudaMalloc(&in, 8192 * sizeof(double));
  cudaMemset(in, 0, 8192 * sizeof(double));
  for (int cfg = 0; cfg < 4; ++cfg) {
    const int threads[4] = {128, 256, 512, 128};
    const int ctas[4] = {1, 1, 1, 4};
    run<0>(threads[cfg], ctas[cfg], "dmma only", out, in);
    run<1>(threads[cfg], ctas[cfg], "dmma + 4 dmul", out, in);
    run<2>(threads[cfg], ctas[cfg], "dmma + lds", out, in);
    run<3>(threads[cfg], ctas[cfg], "dmma + lds + dmul", out, in);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
