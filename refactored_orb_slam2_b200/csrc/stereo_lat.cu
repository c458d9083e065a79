// Latency shape of the stereo matcher (see stereo.cu): 8 left keypoints per CTA.
#define ORBGPU_STEREO_WARPS 8
#define ORBGPU_STEREO_MIN_CTAS 1
#define ORBGPU_VARIANT _lat
#define ORBGPU_VARIANT_IS_SECONDARY 1
#include "stereo.cu"
