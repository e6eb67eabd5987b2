#!/bin/sh
# oracle/ref_build.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own src/ORBextractor.cc, verbatim and from where it lies under
# /root/reference (nothing is copied into the repo), against oracle/cvshim, into
# oracle/_ref/libref_orbextractor.so (git-ignored; travels to the GPU box with gpurun).
# -ffp-contract=off: fp32 expressions evaluated as written (the reference's own
# `-O3 -march=native` build may contract a*b+c into FMA; see DESIGN.md "float parity").
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${ORBFE_REFERENCE:-/root/reference}"
if [ ! -f "$REF/src/ORBextractor.cc" ]; then
    echo "ref_build: $REF/src/ORBextractor.cc not present; keeping any prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$HERE/_ref"
${ORACLE_CXX:-g++} -O2 -std=c++17 -fPIC -ffp-contract=off -w -shared \
    -I"$HERE/cvshim" -I"$REF/include" \
    "$REF/src/ORBextractor.cc" "$HERE/ref_driver.cpp" "$HERE/cvprims.cpp" \
    -o "$HERE/_ref/libref_orbextractor.so"
echo "built $HERE/_ref/libref_orbextractor.so"

# The reference's matcher functions (ORBmatcher::SearchByProjection x3, SearchForInitialization,
# DescriptorDistance, Frame::GetFeaturesInArea / ComputeStereoMatches, MapPoint::PredictScale): their files
# need Eigen/Sophus/DBoW2/g2o as a whole, so ref_slices.py cuts those function bodies out of the tree into a
# temporary translation unit, compiled verbatim against oracle/refshim and deleted afterwards.  The vendored
# DBoW2 (Thirdparty/DBoW2: BowVector, FeatureVector, FORB, ScoringObject, TemplatedVocabulary.h, DUtils/Random, DUtils/Timestamp)
# compiles as whole files; only Boost.Serialization headers and cv::FileStorage are stubbed.
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
python3 "$HERE/ref_slices.py" "$REF" "$TMP/ref_matcher_slices.cc"
${ORACLE_CXX:-g++} -O2 -std=c++17 -fPIC -ffp-contract=off -w -shared \
    -I"$HERE/refshim" -I"$HERE/cvshim" -I"$REF/include" -I"$REF" \
    "$TMP/ref_matcher_slices.cc" "$REF/src/ORBextractor.cc" "$HERE/ref_match_driver.cpp" "$HERE/cvprims.cpp" \
    "$REF/Thirdparty/DBoW2/DBoW2/BowVector.cpp" "$REF/Thirdparty/DBoW2/DBoW2/FeatureVector.cpp" \
    "$REF/Thirdparty/DBoW2/DBoW2/FORB.cpp" "$REF/Thirdparty/DBoW2/DBoW2/ScoringObject.cpp" \
    "$REF/Thirdparty/DBoW2/DUtils/Random.cpp" "$REF/Thirdparty/DBoW2/DUtils/Timestamp.cpp" \
    -o "$HERE/_ref/libref_orbmatcher.so"
echo "built $HERE/_ref/libref_orbmatcher.so"

# KannalaBrandt8::project / unproject (src/CameraModels/KannalaBrandt8.cpp): the file needs Eigen and Boost as a whole;
# the three float bodies are sliced the same way and compiled against refshim/kb8shim.h.
python3 "$HERE/ref_slices.py" "$REF" "$TMP/ref_kb8_slices.cc" kb8
${ORACLE_CXX:-g++} -O2 -std=c++17 -fPIC -ffp-contract=off -w -shared \
    -I"$HERE/refshim" -I"$HERE/cvshim" \
    "$TMP/ref_kb8_slices.cc" "$HERE/ref_kb8_driver.cpp" \
    -o "$HERE/_ref/libref_kb8.so"
echo "built $HERE/_ref/libref_kb8.so"
