// Probe: 3-D TMA box with a 16-byte inner dimension at unaligned / negative inner coordinates (edt_tma.cu's bit tiles).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, unsigned* out)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t tile = smem_u32(smem), bar = tile + 512;
  if (threadIdx.x == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(bar) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(tile),
                 "l"(reinterpret_cast<uint64_t>(&map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
  }
  __syncwarp();
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bar) : "memory");
  for (int i = threadIdx.x; i < 128; i += 32)
    out[i] = reinterpret_cast<unsigned*>(smem)[i];
}
int main()
{
  const int W = 16, NX = 8, NY = 16;
  std::vector<unsigned> h(W * NX * NY);
  for (int x = 0; x < NX; ++x)
    for (int y = 0; y < NY; ++y)
      for (int w = 0; w < W; ++w)
        h[(x * NY + y) * W + w] = 0x1000000u * (x + 1) + 0x1000u * (y + 1) + w + 1;
  unsigned *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, 512);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap m;
  const cuuint64_t dims[3] = { W, NX, NY };
  const cuuint64_t str[2] = { (cuuint64_t)NY * W * 4, (cuuint64_t)W * 4 };
  const cuuint32_t box[3] = { 4, 4, 8 };
  const cuuint32_t es[3] = { 1, 1, 1 };
  CUresult r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  const int tests[][3] = { { 0, 0, 0 }, { 4, 4, 8 }, { 1, 0, 0 }, { -1, 0, 0 }, { 0, 0, -8 }, { 13, 6, 12 } };
  for (auto& t : tests)
  {
    probe<<<1, 32, 1024>>>(m, t[0], t[1], t[2], o);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned res[128];
    cudaMemcpy(res, o, 512, cudaMemcpyDeviceToHost);
    printf("coords (%d,%d,%d): %s  first row:", t[0], t[1], t[2], cudaGetErrorString(e));
    for (int i = 0; i < 8; ++i)
      printf(" %08x", res[i]);
    printf(" ... [y1] %08x\n", res[16]);
    if (e != cudaSuccess)
      break;
  }
  return 0;
}
