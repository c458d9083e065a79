// The CUDA programming guide's own 2-D TMA example (libcu++ wrappers), int32 64x64 tile.
#include <cstdio>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int GW = 1024, GH = 1024, SW = 32, SH = 32;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out) {
  __shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < SH * SW; i += blockDim.x) out[i] = smem_buffer[i / SW][i % SW];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  std::vector<int> h(GW * GH); for (int i = 0; i < GW * GH; i++) h[i] = i * 7 + 1;
  int *d, *out; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice); cudaMalloc(&out, SW * SH * 4);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap map{};
  cuuint64_t size[2] = {GW, GH}; cuuint64_t stride[1] = {GW * sizeof(int)}; cuuint32_t box[2] = {SW, SH}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  kernel<<<1, 128>>>(map, 64, 96, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("guide example: %s", cudaGetErrorString(e));
  if (!e) { std::vector<int> o(SW * SH); cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost); int bad = 0;
    for (int y = 0; y < SH; y++) for (int x = 0; x < SW; x++) bad += o[y * SW + x] != h[(96 + y) * GW + 64 + x]; printf(" mismatches %d", bad); }
  printf("\n");
  int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt); printf("driver %d runtime %d\n", drv, rt);
  return 0;
}
