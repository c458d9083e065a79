// Latency shape of the octree kernel (see octree.cu): 512 threads per (image, level) CTA.
#define ORBGPU_OCTREE_THREADS 512
#define ORBGPU_VARIANT _lat
#define ORBGPU_VARIANT_IS_SECONDARY 1
#include "octree.cu"
