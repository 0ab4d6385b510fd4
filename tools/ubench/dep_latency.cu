// dep_latency.cu -- what does a LONE warp pay per dependent FP32 instruction on sm_100a?
// (background for chain_team_kernel: the SVF recurrence is a 4-deep chain FADD -> FMUL -> FADD -> FFMA per sample)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dep_latency dep_latency.cu && ./dep_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, float a, float b, int n)
{
  float x = a + threadIdx.x, y = b, ic1 = 0.1f, ic2 = 0.2f;
  const float g0 = 0.3f, g1 = -0.2f, g2 = 0.05f;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i)
  {
    if (MODE == 0) { x = __fadd_rn(x, y); }                                  // FADD chain
    if (MODE == 1) { x = __fmul_rn(x, y); }                                  // FMUL chain
    if (MODE == 2) { x = __fmaf_rn(x, y, b); }                               // FFMA chain
    if (MODE == 3) { x = __fadd_rn(__fmul_rn(x, y), b); }                    // FMUL -> FADD
    if (MODE == 4)
    {  // the exact SVF sample: chain through ic2 is FADD -> FMUL -> FADD -> FFMA
      float t = __fsub_rn(x, ic2);
      float t1 = __fadd_rn(__fmul_rn(g0, t), __fmul_rn(g1, ic1));
      float t2 = __fadd_rn(__fmul_rn(g2, t), __fmul_rn(g0, ic1));
      y = __fadd_rn(t2, ic2);
      ic1 = __fmaf_rn(2.f, t1, ic1);
      ic2 = __fmaf_rn(2.f, t2, ic2);
    }
    if (MODE == 5)
    {  // the same with FMA contraction allowed (fast mode): FADD -> FFMA -> FFMA
      float t = x - ic2;
      float t1 = fmaf(g0, t, g1 * ic1);
      float t2 = fmaf(g2, t, g0 * ic1);
      y = t2 + ic2;
      ic1 = fmaf(2.f, t1, ic1);
      ic2 = fmaf(2.f, t2, ic2);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + ic1 + ic2;
}

int main()
{
  float* d;
  long long* c;
  cudaMalloc(&d, 1 << 20);
  cudaMalloc(&c, 8);
  const int n = 1 << 16;
  const char* names[] = {"FADD chain", "FMUL chain", "FFMA chain", "FMUL->FADD", "SVF sample exact", "SVF sample fma"};
  for (int warps = 1; warps <= 4; warps *= 2)
    for (int m = 0; m < 6; ++m)
    {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep)
      {
        if (m == 0) k<0><<<1, 32 * warps * 4>>>(d, c, 1.f, 1e-9f, n);
        if (m == 1) k<1><<<1, 32 * warps * 4>>>(d, c, 1.f, 1.0000001f, n);
        if (m == 2) k<2><<<1, 32 * warps * 4>>>(d, c, 1.f, 0.999f, n);
        if (m == 3) k<3><<<1, 32 * warps * 4>>>(d, c, 1.f, 0.999f, n);
        if (m == 4) k<4><<<1, 32 * warps * 4>>>(d, c, 1.f, 0.999f, n);
        if (m == 5) k<5><<<1, 32 * warps * 4>>>(d, c, 1.f, 0.999f, n);
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      }
      printf("%d warp(s)/scheduler  %-18s %6.2f cycles per iteration\n", warps, names[m], (double)h / n);
    }
  // effective SM clock of a short, light kernel: cycles counted by clock64 / CUDA-event time
  {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep)
    {
      long long h = 0;
      float ms = 0;
      cudaEventRecord(e0);
      k<4><<<148, 128>>>(d, c, 1.f, 0.999f, 1 << 14);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      printf("light kernel: %lld cycles in %.3f us -> %.0f MHz effective SM clock\n", h, ms * 1e3, (double)h / (ms * 1e3));
    }
  }
  return 0;
}
