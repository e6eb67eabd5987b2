#!/usr/bin/env python3
"""oracle/ref_slices.py -- TEST INFRASTRUCTURE ONLY.

Writes ONE translation unit made of function bodies cut, at build time, out of the reference's own
sources where they lie under /root/reference (src/ORBmatcher.cc, src/Frame.cc, src/MapPoint.cc).  The
whole files cannot be compiled here (Eigen, Sophus' Eigen dependency, DBoW2's boost, g2o are absent),
but the matcher functions on the hot path only touch a handful of members, which oracle/refshim
declares.  The output goes to a temporary directory chosen by oracle/ref_build.sh and is deleted after
compilation: no reference source text is ever stored in this repository.

usage: ref_slices.py <reference root> <output .cc> [kb8]
  kb8: the second, independent unit -- KannalaBrandt8::project (both float overloads) and ::unproject out of
       src/CameraModels/KannalaBrandt8.cpp, compiled against oracle/refshim/kb8shim.h.
"""
import re
import sys

# (file, regex matching the start of the definition's first line)
SLICES = [
    ("src/ORBmatcher.cc", r"const int ORBmatcher::TH_HIGH\s*="),
    ("src/ORBmatcher.cc", r"const int ORBmatcher::TH_LOW\s*="),
    ("src/ORBmatcher.cc", r"const int ORBmatcher::HISTO_LENGTH\s*="),
    ("src/ORBmatcher.cc", r"ORBmatcher::ORBmatcher\("),
    ("src/ORBmatcher.cc", r"int ORBmatcher::SearchByProjection\(Frame &F, const vector<MapPoint\*> &vpMapPoints"),
    ("src/ORBmatcher.cc", r"float ORBmatcher::RadiusByViewingCos\("),
    ("src/ORBmatcher.cc", r"int ORBmatcher::SearchByBoW\(KeyFrame\* pKF,Frame &F"),
    ("src/ORBmatcher.cc", r"int ORBmatcher::SearchByBoW\(KeyFrame \*pKF1, KeyFrame \*pKF2"),
    ("src/ORBmatcher.cc", r"int ORBmatcher::SearchForTriangulation\("),
    ("src/CameraModels/Pinhole.cpp", r"bool Pinhole::epipolarConstrain\("),
    ("src/ORBmatcher.cc", r"int ORBmatcher::SearchForInitialization\("),
    ("src/ORBmatcher.cc", r"int ORBmatcher::SearchByProjection\(Frame &CurrentFrame, const Frame &LastFrame"),
    ("src/ORBmatcher.cc", r"int ORBmatcher::SearchByProjection\(Frame &CurrentFrame, KeyFrame \*pKF"),
    ("src/ORBmatcher.cc", r"int ORBmatcher::SearchByProjection\(KeyFrame\* pKF, Sophus::Sim3f &Scw, const vector<MapPoint\*> &vpPoints,\s*vector<MapPoint\*> &vpMatched"),
    ("src/ORBmatcher.cc", r"int ORBmatcher::Fuse\(KeyFrame \*pKF, const vector<MapPoint \*> &vpMapPoints"),
    ("src/ORBmatcher.cc", r"int ORBmatcher::Fuse\(KeyFrame \*pKF, Sophus::Sim3f &Scw"),
    ("src/ORBmatcher.cc", r"int ORBmatcher::SearchBySim3\("),
    ("src/ORBmatcher.cc", r"void ORBmatcher::ComputeThreeMaxima\("),
    ("src/ORBmatcher.cc", r"int ORBmatcher::DescriptorDistance\("),
    ("src/Frame.cc", r"void Frame::AssignFeaturesToGrid\("),
    ("src/Frame.cc", r"vector<size_t> Frame::GetFeaturesInArea\("),
    ("src/Frame.cc", r"bool Frame::PosInGrid\("),
    ("src/Frame.cc", r"void Frame::ComputeStereoMatches\("),
    ("src/MapPoint.cc", r"int MapPoint::PredictScale\(const float &currentDist, Frame\* pF\)"),
    ("src/MapPoint.cc", r"int MapPoint::PredictScale\(const float &currentDist, KeyFrame\* pKF\)"),
    ("src/MapPoint.cc", r"void MapPoint::ComputeDistinctiveDescriptors\("),
    ("src/KeyFrame.cc", r"vector<size_t> KeyFrame::GetFeaturesInArea\("),
    ("src/KeyFrame.cc", r"bool KeyFrame::IsInImage\("),
]


KB8_SLICES = [
    ("src/CameraModels/KannalaBrandt8.cpp", r"cv::Point2f KannalaBrandt8::project\(const cv::Point3f &p3D\)"),
    ("src/CameraModels/KannalaBrandt8.cpp", r"Eigen::Vector2f KannalaBrandt8::project\(const Eigen::Vector3f &v3D\)"),
    ("src/CameraModels/KannalaBrandt8.cpp", r"cv::Point3f KannalaBrandt8::unproject\(const cv::Point2f &p2D\)"),
]


def blank_comments(text):
    """Same length as `text`, with comments and string/char literals replaced by spaces."""
    out = list(text)
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if text.startswith("//", i):
            j = text.find("\n", i)
            j = n if j < 0 else j
            out[i:j] = " " * (j - i)
            i = j
        elif text.startswith("/*", i):
            j = text.find("*/", i + 2)
            j = n if j < 0 else j + 2
            for k in range(i, j):
                if out[k] != "\n":
                    out[k] = " "
            i = j
        elif c in "\"'":
            j = i + 1
            while j < n and text[j] != c:
                j += 2 if text[j] == "\\" else 1
            for k in range(i + 1, min(j, n)):
                out[k] = " "
            i = j + 1
        else:
            i += 1
    return "".join(out)


def cut(text, code, pattern):
    m = re.search(r"^[ \t]*" + pattern, code, re.M)
    if not m:
        raise SystemExit("ref_slices: no definition matching %r" % pattern)
    start = m.start()
    semi = code.find(";", m.end())
    brace = code.find("{", m.end())
    if brace < 0 or (0 <= semi < brace):  # a plain statement (constant definition)
        return text[start : semi + 1], text.count("\n", 0, start) + 1
    depth, i = 0, brace
    while True:
        if code[i] == "{":
            depth += 1
        elif code[i] == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    return text[start : i + 1], text.count("\n", 0, start) + 1


def main(ref, out_path, mode=""):
    cache = {}
    parts = [
        "// GENERATED by oracle/ref_slices.py from the reference tree; temporary, never committed.\n",
        '#include "refshim.h"\n#include "ORBmatcher.h"\n#include <limits.h>\n#include <stdint.h>\nusing namespace std;\nnamespace ORB_SLAM3 {\n',
    ]
    if mode == "kb8":   # the reference file has no `using namespace std`
        parts[1] = '#include "kb8shim.h"\nnamespace ORB_SLAM3 {\n'
    for rel, pat in (KB8_SLICES if mode == "kb8" else SLICES):
        if rel not in cache:
            t = open("%s/%s" % (ref, rel), encoding="utf-8", errors="replace").read()
            cache[rel] = (t, blank_comments(t))
        t, code = cache[rel]
        body, line = cut(t, code, pat)
        parts.append('\n#line %d "%s/%s"\n%s\n' % (line, ref, rel, body))
    parts.append("\n}  // namespace ORB_SLAM3\n")
    open(out_path, "w", encoding="utf-8").write("".join(parts))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
