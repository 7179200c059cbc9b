// Exhaustive error check of the FAST kernel's sine / cosine (validity_fast.cu: sincos_small) over every fp32 value in
// [-kAngleRange, kAngleRange]: max error in ulps of 1.0 (2^-23) against fp64 sin / cos.  The FAST path's error analysis
// budgets 2 such ulps per value (validity_fast.cu::fast_create, "sc").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sincos_check tools/sincos_check.cu && ./sincos_check
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#include "../moveit_b200/csrc/sincos_small.cuh"

__global__ void check(uint32_t first_bits, uint32_t count, double* max_err)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double worst = 0.0;
  for (uint32_t k = i; k < count; k += gridDim.x * blockDim.x)
  {
    const uint32_t bits = first_bits + k;
    for (int sign = 0; sign < 2; ++sign)
    {
      const float x = __uint_as_float(bits | (sign ? 0x80000000u : 0u));
      float s, c;
      b200df::sincos_small(x, &s, &c);
      const double es = fabs((double)s - sin((double)x)), ec = fabs((double)c - cos((double)x));
      worst = fmax(worst, fmax(es, ec));
    }
  }
  // block max -> global max (double compare via atomicMax on the bit pattern: non-negative doubles order like integers)
  atomicMax(reinterpret_cast<unsigned long long*>(max_err), (unsigned long long)__double_as_longlong(worst));
}

int main()
{
  double* d;
  cudaMalloc(&d, sizeof(double));
  cudaMemset(d, 0, sizeof(double));
  float hi = 8.0f;
  uint32_t hi_bits;
  std::memcpy(&hi_bits, &hi, 4);
  check<<<148 * 8, 256>>>(0u, hi_bits + 1u, d);  // every non-negative float up to 8.0 (and its negative)
  double h = 0.0;
  cudaMemcpy(&h, d, sizeof(double), cudaMemcpyDeviceToHost);
  const double ulp1 = 1.0 / 8388608.0;
  std::printf("max |error| over [-8, 8]: %.3e = %.3f ulp(1.0)\n", h, h / ulp1);
  return h / ulp1 <= 2.0 ? 0 : 1;
}
