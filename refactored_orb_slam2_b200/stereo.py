"""Frame::ComputeStereoMatches (Source/Libraries/ORB_SLAM2/src/Frame.cc:477-646) over the C ABI."""
import numpy as np

from ._lib import check, lib, ptr


def ComputeStereoMatches(extractor_left, extractor_right, mbf, mb, first_left=0, first_right=0, pairs=1):
    """Runs on the device-resident keypoints / descriptors / pyramids of the two extractors' last
    calls.  `mb` is the stereo baseline in metres (= mbf / fx); the reference reads that member
    before assigning it (Frame.cc:505 vs :124), so it is an explicit argument here.
    Returns (mvuRight, mvDepth): float32 [pairs, capacity], -1 where a keypoint has no match."""
    cap = extractor_left.capacity
    u = np.empty((pairs, cap), np.float32)
    z = np.empty((pairs, cap), np.float32)
    check(lib().orbgpu_stereo_match(extractor_left.handle, extractor_right.handle, first_left, first_right, pairs,
                                    float(mbf), float(mb), ptr(u), ptr(z), cap))
    return u, z


class Frame:
    """The stereo constructor's front-end steps only (Frame.cc:66-127): two extractions and
    ComputeStereoMatches.  Members use the reference's names."""

    def __init__(self, imLeft, imRight, extractorLeft, extractorRight, bf, fx):
        self.mbf = np.float32(bf)
        self.mb = np.float32(self.mbf / np.float32(fx))      # Frame.cc:124
        self.mvKeys, self.mDescriptors = extractorLeft(imLeft)
        self.mvKeysRight, self.mDescriptorsRight = extractorRight(imRight)
        self.N = len(self.mvKeys)
        if self.N and len(self.mvKeysRight):
            u, z = ComputeStereoMatches(extractorLeft, extractorRight, self.mbf, self.mb)
            self.mvuRight, self.mvDepth = u[0, :self.N].copy(), z[0, :self.N].copy()
        else:
            self.mvuRight = np.full(self.N, -1, np.float32)
            self.mvDepth = np.full(self.N, -1, np.float32)
