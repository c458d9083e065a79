// Step-by-step departures from the guide's working example, u8 tensor.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int GW = 1280, GH = 414;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, unsigned char* out, int bytes) {
  __shared__ alignas(128) unsigned char smem_buffer[160 * 64];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem_buffer[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  int boxW = atoi(argv[1]), boxH = atoi(argv[2]), x = atoi(argv[3]), y = atoi(argv[4]);
  std::vector<unsigned char> h(GW * GH); for (int i = 0; i < GW * GH; i++) h[i] = (unsigned char)((i * 2654435761u) >> 13);
  unsigned char *d, *out; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice); cudaMalloc(&out, 160 * 64);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap map{};
  cuuint64_t size[2] = {GW, GH}; cuuint64_t stride[1] = {GW}; cuuint32_t box[2] = {(cuuint32_t)boxW, (cuuint32_t)boxH}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  kernel<<<1, 128>>>(map, x, y, out, boxW * boxH);
  cudaError_t e = cudaDeviceSynchronize();
  printf("u8 box %dx%d at (%d,%d) encode %d: %s", boxW, boxH, x, y, (int)r, cudaGetErrorString(e));
  if (!e) { std::vector<unsigned char> o(boxW * boxH); cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost); int bad = 0;
    for (int yy = 0; yy < boxH; yy++) for (int xx = 0; xx < boxW; xx++) { int gx = x + xx, gy = y + yy; unsigned char ex = (gx >= 0 && gx < GW && gy >= 0 && gy < GH) ? h[gy * GW + gx] : 0; bad += o[yy * boxW + xx] != ex; } printf(" mismatches %d", bad); }
  printf("\n");
  return 0;
}
