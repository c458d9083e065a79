// Exercises the shim exactly the way Frame's stereo constructor does (Frame.cc:66-127): two
// extractor objects, two operator() calls, then stereo matching.  Input: two raw 8-bit images.
// Output: a binary dump the Python parity test compares with the ctypes path.
//   shim_demo left.raw right.raw W H nfeatures out.bin [repeats]
// With `repeats` it also times the same three calls (the reference's per-frame calling pattern through the C++ drop-in a
// maintainer links, not through ctypes) with the host pyramid download of mvImagePyramid on and off, and prints one JSON line.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "FrameStereo.h"
#include "ORBextractor.h"

static std::vector<uint8_t> ReadAll(const char *path, size_t n) {
  std::vector<uint8_t> v(n);
  FILE *f = fopen(path, "rb");
  if (!f || fread(v.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return v;
}

int main(int argc, char **argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s left.raw right.raw W H nfeatures out.bin\n", argv[0]); return 2; }
  const int W = atoi(argv[3]), H = atoi(argv[4]), nf = atoi(argv[5]);
  std::vector<uint8_t> l = ReadAll(argv[1], (size_t)W * H), r = ReadAll(argv[2], (size_t)W * H);
  cv::Mat imLeft(H, W, CV_8UC1, l.data()), imRight(H, W, CV_8UC1, r.data());
  try {
    ORB_SLAM2::ORBextractor exL(nf, 1.2f, 8, 20, 7), exR(nf, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> kL, kR;
    cv::Mat dL, dR;
    exL(imLeft, cv::Mat(), kL, dL);
    exR(imRight, cv::Mat(), kR, dR);
    std::vector<float> uR, depth;
    const float bf = 386.1448f, fx = 718.856f;
    ORB_SLAM2::ComputeStereoMatchesGPU(exL, exR, bf, bf / fx, (int)kL.size(), uR, depth);
    FILE *f = fopen(argv[6], "wb");
    const int n = (int)kL.size(), levels = exL.GetLevels();
    fwrite(&n, 4, 1, f);
    fwrite(kL.data(), sizeof(cv::KeyPoint), n, f);
    for (int i = 0; i < n; i++) fwrite(dL.ptr(i), 1, 32, f);
    fwrite(uR.data(), 4, n, f);
    fwrite(depth.data(), 4, n, f);
    fwrite(&levels, 4, 1, f);
    const cv::Mat &top = exL.mvImagePyramid[levels - 1];
    fwrite(&top.cols, 4, 1, f); fwrite(&top.rows, 4, 1, f);
    for (int y = 0; y < top.rows; y++) fwrite(top.ptr(y), 1, top.cols, f);
    fclose(f);
    printf("shim_demo: %d keypoints, scale[1] = %.7f\n", n, exL.GetScaleFactors()[1]);
    const int repeats = argc > 7 ? atoi(argv[7]) : 0;
    if (repeats > 0) {
      double med[2] = {0, 0};
      for (int mode = 0; mode < 2; mode++) {                  // 0: mvImagePyramid downloaded every call (default), 1: off
        exL.SetPyramidDownload(mode == 0); exR.SetPyramidDownload(mode == 0);
        std::vector<double> ms;
        for (int i = 0; i < repeats + 5; i++) {
          const auto t0 = std::chrono::steady_clock::now();
          std::thread right([&] { exR(imRight, cv::Mat(), kR, dR); });      // Frame.cc:87-99: the two extractions run on two threads
          exL(imLeft, cv::Mat(), kL, dL);
          right.join();
          ORB_SLAM2::ComputeStereoMatchesGPU(exL, exR, bf, bf / fx, (int)kL.size(), uR, depth);
          const auto t1 = std::chrono::steady_clock::now();
          if (i >= 5) ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
        }
        std::sort(ms.begin(), ms.end());
        med[mode] = ms[ms.size() / 2];
      }
      printf("{\"what\": \"C++ drop-in, one %dx%d stereo frame: ORBextractor::operator() on two threads (Frame.cc:87-99) + ComputeStereoMatchesGPU, pageable host buffers\", "
             "\"keypoints\": %d, \"repeats\": %d, \"ms_with_pyramid_download\": %.4f, \"ms_without_pyramid_download\": %.4f}\n",
             W, H, n, repeats, med[0], med[1]);
    }
  } catch (const std::exception &e) {
    fprintf(stderr, "shim_demo: %s\n", e.what());
    return 1;
  }
  return 0;
}
