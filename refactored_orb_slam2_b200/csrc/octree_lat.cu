// Latency shape of the octree kernel (see octree.cu): 1024 threads per (image, level) CTA (0.265 ms per stereo pair against 0.278 at 512).
#define ORBGPU_OCTREE_THREADS 1024
#define ORBGPU_OCTREE_MIN_CTAS 1
#define ORBGPU_VARIANT _lat
#define ORBGPU_VARIANT_IS_SECONDARY 1
#include "octree.cu"
