"""orbfe -- host-side mirror of the reference's ORB front-end interface on top of the C ABI of
libORBfe_b200.so (CUDA, sm_100a).

  ORBextractor  <->  ORB_SLAM3::ORBextractor   (/root/reference/include/ORBextractor.h:46-112)
  ORBmatcher    <->  ORB_SLAM3::ORBmatcher     (/root/reference/include/ORBmatcher.h:33-104), the
                     Frame-based hot-path subset: DescriptorDistance, the three Frame overloads of
                     SearchByProjection, Frame::ComputeStereoMatches and the kNN-2 + ratio matcher of
                     Frame::ComputeStereoFishEyeMatches; plus the SURVEY 8(f) "next" searches
                     (SearchForInitialization, Fuse, SearchBySim3, Sim3 SearchByProjection, SearchByBoW).
  ORBVocabulary <->  ORB_SLAM3::ORBVocabulary  (/root/reference/include/ORBVocabulary.h:30-31, DBoW2 transform)
  KannalaBrandt8 <-> ORB_SLAM3::KannalaBrandt8 (/root/reference/include/CameraModels/KannalaBrandt8.h): project,
                     unproject, TriangulateMatches / epipolarConstrain, batched over points / matches

The reference is C++; the drop-in binding for it is the header-only adapter under host/.  This
Python mirror exists so that tests/, bench.py and multi-GPU drivers exercise exactly the same C
entry points.  No function here computes on the CPU (ORBVocabulary.transform folds the per-feature
device results into the BowVector / FeatureVector maps on the host, as the C++ adapter does)."""
from ._lib import (EMPTY_IMAGE, ERR_CAPACITY, ERR_CUDA, ERR_INVALID, EXPORTS, KP_DTYPE, LIB_PATH, OK,
                   STAGE_NAMES, OrbfeError, last_error, lib)
from .extractor import ORBextractor
from .matcher import FrameData, ORBmatcher
from .vocabulary import ORBVocabulary
from . import intake
from .kb8 import KannalaBrandt8

__all__ = ["ORBextractor", "ORBmatcher", "ORBVocabulary", "KannalaBrandt8", "intake", "FrameData", "KP_DTYPE", "OrbfeError", "lib", "LIB_PATH", "EXPORTS"]
