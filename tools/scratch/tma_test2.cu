#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int bytes, unsigned char* out, int fenceKind) {
  __shared__ __align__(1024) unsigned char s[144 * 64];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    if (fenceKind == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(s)), "l"(&map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(s)), "l"(&map), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int pitch = 1280, rows = 414, batch = 3; const size_t plane = (size_t)pitch * rows;
  std::vector<unsigned char> h(plane * batch);
  for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)((i * 2654435761u) >> 13);
  unsigned char *d, *out; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  int boxW = (variant == 2) ? 128 : 144, boxH = 38; cudaMalloc(&out, 144 * 64);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("variant %d entry %s q=%d p=%p\n", variant, cudaGetErrorString(ge), (int)q, p);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  alignas(64) CUtensorMap map;
  int rank = (variant == 1) ? 2 : 3;
  cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows * (rank == 2 ? batch : 1), (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)plane};
  cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  variant == 3 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode result %d\n", (int)r);
  const unsigned long long* w = (const unsigned long long*)&map;
  for (int i = 0; i < 16; i++) printf("%016llx%s", w[i], i % 4 == 3 ? "\n" : " ");
  int c0 = 34, c1 = 35, c2 = 1, bytes = boxW * boxH;
  if (rank == 3) k<3><<<1, 128>>>(map, c0, c1, c2, bytes, out, variant == 4 ? 1 : 0);
  else k<2><<<1, 128>>>(map, c0, c1 + rows * c2, 0, bytes, out, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<unsigned char> o(bytes); cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < boxH; y++) for (int x = 0; x < boxW; x++) bad += o[y * boxW + x] != h[(size_t)c2 * plane + (size_t)(c1 + y) * pitch + c0 + x];
    printf("  mismatches %d", bad);
  }
  printf("\n");
  return 0;
}
