"""orbfe -- host-side mirror of the reference's ORB front-end interface on top of the C ABI of
libORBfe_b200.so (CUDA, sm_100a).

  ORBextractor  <->  ORB_SLAM3::ORBextractor   (/root/reference/include/ORBextractor.h:46-112)
  ORBmatcher    <->  ORB_SLAM3::ORBmatcher     (/root/reference/include/ORBmatcher.h:33-104), the
                     Frame-based hot-path subset: DescriptorDistance, the three Frame overloads of
                     SearchByProjection, Frame::ComputeStereoMatches and the kNN-2 + ratio matcher of
                     Frame::ComputeStereoFishEyeMatches.

The reference is C++; the drop-in binding for it is the header-only adapter under host/.  This
Python mirror exists so that tests/, bench.py and multi-GPU drivers exercise exactly the same C
entry points.  No function here computes on the CPU."""
from ._lib import (EMPTY_IMAGE, ERR_CAPACITY, ERR_CUDA, ERR_INVALID, EXPORTS, KP_DTYPE, LIB_PATH, OK,
                   STAGE_NAMES, OrbfeError, last_error, lib)
from .extractor import ORBextractor
from .matcher import FrameData, ORBmatcher

__all__ = ["ORBextractor", "ORBmatcher", "FrameData", "KP_DTYPE", "OrbfeError", "lib", "LIB_PATH", "EXPORTS"]
