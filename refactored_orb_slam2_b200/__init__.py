"""refactored_orb_slam2_b200 — the ORB front-end of sjulier/Refactored_ORB_SLAM2 as hand-written
sm_100a CUDA kernels behind a C ABI (include/orbgpu.h) and a host-side mirror of the reference's
ORBextractor / ORBmatcher / Frame::ComputeStereoMatches interfaces.  No CPU fallback."""
from ._lib import KP_DTYPE, TOP2_DTYPE, Camera, OrbGpuError, lib  # noqa: F401
from .extractor import ORBextractor  # noqa: F401
from .features import ComputeImageBounds, FrameFeatures, UndistortKeyPoints, grid_geometry, make_policy, undistort_points  # noqa: F401
from .matcher import DescriptorBank, ORBmatcher, distinctive_descriptors, merge_top2_device  # noqa: F401
from .stereo import ComputeStereoMatches, Frame  # noqa: F401
from .pipeline import StereoBatchPipeline  # noqa: F401
from .vocabulary import Vocabulary  # noqa: F401
from .bowdb import BowDatabase  # noqa: F401
from . import staging  # noqa: F401

__all__ = ["ORBextractor", "FrameFeatures", "Camera", "UndistortKeyPoints", "undistort_points", "ComputeImageBounds", "grid_geometry", "make_policy", "Vocabulary", "BowDatabase", "distinctive_descriptors", "ORBmatcher", "DescriptorBank", "ComputeStereoMatches", "Frame", "StereoBatchPipeline", "merge_top2_device",
           "KP_DTYPE", "TOP2_DTYPE", "OrbGpuError", "lib"]
