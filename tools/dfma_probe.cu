// DFMA issue-rate probe: does the operand pattern limit the fp64 pipe?
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(const double* __restrict__ in, double* out, int iters) {
  double a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 7]; c[i] = in[threadIdx.x + i]; }
  const double u = in[blockIdx.x & 1];  // warp-uniform at run time
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) c[i] = fma(c[i], 0.999999, 1e-7);            // reg, imm/const, imm/const
        if (MODE == 1) c[i] = fma(a[i], b[i], c[i]);                 // 3 distinct register operands
        if (MODE == 2) c[i] = fma(a[0], b[i], c[i]);                 // one operand repeated (reuse cache)
        if (MODE == 3) c[i] = fma(c[i], u, a[i]);                    // reg, reg(uniform value), reg
        if (MODE == 4) c[i] = fma(a[i], a[i], c[i]);                 // 2 distinct registers
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i];
  if (s == 1.2345) out[0] = s;
}
template <int MODE>
void run(const double* in, double* out, const char* what) {
  int dev = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ctas_per_sm : {2, 4, 8}) {
    k<MODE><<<sms * ctas_per_sm, 256>>>(in, out, iters / 10);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0); k<MODE><<<sms * ctas_per_sm, 256>>>(in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = 2.0 * 32.0 * iters * (double)sms * ctas_per_sm * 256 / (best * 1e-3);
    printf("%-44s CTAs/SM %d: %.2f TFLOP/s\n", what, ctas_per_sm, flops / 1e12);
  }
}
int main() {
  double *in, *out; cudaMalloc(&in, 4096 * 8); cudaMalloc(&out, 8);
  double h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1e-3 * (i % 17) + 0.5; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>(in, out, "c = fma(c, const, const)");
  run<1>(in, out, "c = fma(a[i], b[i], c)  3 registers");
  run<2>(in, out, "c = fma(a[0], b[i], c)  one repeated");
  run<3>(in, out, "c = fma(c, u, a[i])");
  run<4>(in, out, "c = fma(a[i], a[i], c)  2 registers");
  return 0;
}
