// DMMA probe: does mma.sync.m8n8k4.f64 issue on a pipe separate from DFMA on sm_100a, and at what
// rate alone / next to a DFMA stream?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// MODE 0: every warp DMMA; 1: every warp DFMA (2-register form: full rate); 2: even warps DMMA, odd warps DFMA;
// 3: every warp interleaves DMMA and DFMA in one instruction stream
template <int MODE>
__global__ void __launch_bounds__(256) k(const double* __restrict__ in, double* out, int iters) {
  const int warp = threadIdx.x >> 5;
  double a[4], b[4], c0[8], c1[8], f[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 7 + i]; }
#pragma unroll
  for (int i = 0; i < 8; ++i) { c0[i] = in[i]; c1[i] = in[i + 9]; f[i] = in[threadIdx.x + 2 * i]; }
  const bool do_mma = MODE == 0 || MODE == 3 || (MODE == 2 && (warp & 1) == 0);
  const bool do_fma = MODE == 1 || MODE == 3 || (MODE == 2 && (warp & 1) == 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (do_mma) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(c0[i], c1[i], a[i & 3], b[(i + r) & 3]);
      }
      if (do_fma) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fma(f[i], f[i], 0.25);  // 1 register: never operand-bandwidth limited
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i] + f[i];
  if (s == 1.2345) out[0] = s;
}

template <int MODE>
void run(const double* in, double* out, const char* what) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int ctas : {1, 2, 4}) {
    k<MODE><<<sms * ctas, 256>>>(in, out, iters / 10);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k<MODE><<<sms * ctas, 256>>>(in, out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double warps = (double)sms * ctas * 8;
    const double mma_warps = MODE == 0 || MODE == 3 ? warps : (MODE == 2 ? warps / 2 : 0);
    const double fma_warps = MODE == 1 || MODE == 3 ? warps : (MODE == 2 ? warps / 2 : 0);
    const double mma_flops = mma_warps * iters * 32.0 * 512.0 / (best * 1e-3);       // 8x8x4 MACs x 2 per DMMA
    const double fma_flops = fma_warps * iters * 32.0 * 32.0 * 2.0 / (best * 1e-3);  // 32 lanes x 2 per DFMA
    printf("%-46s CTAs/SM %d: %7.3f ms  DMMA %6.2f TFLOP/s  DFMA %6.2f TFLOP/s  sum %6.2f\n", what, ctas, best,
           mma_flops / 1e12, fma_flops / 1e12, (mma_flops + fma_flops) / 1e12);
  }
}

int main() {
  double *in, *out;
  cudaMalloc(&in, 4096 * 8);
  cudaMalloc(&out, 8);
  double h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = 1e-3 * (i % 17) + 0.5;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>(in, out, "DMMA m8n8k4 f64, all warps");
  run<1>(in, out, "DFMA, all warps");
  run<2>(in, out, "even warps DMMA, odd warps DFMA");
  run<3>(in, out, "DMMA and DFMA interleaved in every warp");
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
