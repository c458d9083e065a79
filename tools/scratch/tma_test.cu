// Standalone TMA sanity test: loads a (boxW x boxH) u8 window of a [batch][rows][pitch] tensor.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../refactored_orb_slam2_b200/csrc/tma.cuh"
using namespace orbgpu;

__global__ void k_single(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int bytes, unsigned char* out) {
  extern __shared__ unsigned char raw[];
  __shared__ __align__(8) unsigned long long bar;
  unsigned char* s = raw + ((128u - (smem_u32(raw) & 127u)) & 127u);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(&bar, bytes); tma_load_3d(s, &map, c0, c1, c2, &bar); }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = s[i];
}
__global__ void k_indexed(const __grid_constant__ TmaMaps maps, int which, int c0, int c1, int c2, int bytes, unsigned char* out) {
  extern __shared__ unsigned char raw[];
  __shared__ __align__(8) unsigned long long bar;
  unsigned char* s = raw + ((128u - (smem_u32(raw) & 127u)) & 127u);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(&bar, bytes); tma_load_3d(s, &maps.m[which], c0, c1, c2, &bar); }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = s[i];
}
int main() {
  const int pitch = 1280, rows = 414, batch = 3; const size_t plane = (size_t)pitch * rows;  // 529920, multiple of 256
  std::vector<unsigned char> h(plane * batch);
  for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)((i * 2654435761u) >> 13);
  unsigned char *d, *out; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  const int boxW = 144, boxH = 38; cudaMalloc(&out, boxW * boxH);
  TmaMaps maps; 
  for (int i = 0; i < 8; i++) { cudaError_t e = encode_plane_map(&maps.m[i], d, pitch, rows, plane, batch, boxW, boxH); if (e) { printf("encode %d failed %s\n", i, cudaGetErrorString(e)); return 1; } }
  auto check = [&](const char* name, int c0, int c1, int c2) {
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s: %s", name, cudaGetErrorString(e));
    if (e == cudaSuccess) {
      std::vector<unsigned char> o(boxW * boxH); cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int y = 0; y < boxH; y++) for (int x = 0; x < boxW; x++) {
        int gx = c0 + x, gy = c1 + y; unsigned char exp = (gx >= 0 && gx < pitch && gy >= 0 && gy < rows) ? h[(size_t)c2 * plane + (size_t)gy * pitch + gx] : 0;
        bad += o[y * boxW + x] != exp;
      }
      printf("  mismatches %d", bad);
    }
    printf("\n");
  };
  cudaFuncSetAttribute(k_single, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  cudaFuncSetAttribute(k_indexed, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  k_single<<<1, 128, boxW * boxH + 256>>>(maps.m[0], 34, 35, 1, boxW * boxH, out); check("single", 34, 35, 1);
  k_indexed<<<1, 128, boxW * boxH + 256>>>(maps, 3, 158, 66, 2, boxW * boxH, out); check("indexed", 158, 66, 2);
  k_indexed<<<1, 128, boxW * boxH + 256>>>(maps, 5, 1200, 400, 0, boxW * boxH, out); check("indexed-oob", 1200, 400, 0);
  k_indexed<<<1, 128, boxW * boxH + 256>>>(maps, 1, -3, -2, 0, boxW * boxH, out); check("indexed-neg", -3, -2, 0);
  return 0;
}
