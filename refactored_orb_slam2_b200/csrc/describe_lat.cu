// Latency shape of the orientation + descriptor kernel (see describe.cu): 8 keypoints (4 warps) per CTA.
#define ORBGPU_DESCRIBE_WARPS 4
#define ORBGPU_VARIANT _lat
#define ORBGPU_VARIANT_IS_SECONDARY 1
#include "describe.cu"
