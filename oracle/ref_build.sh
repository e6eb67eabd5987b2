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
${CXX:-g++} -O2 -std=c++17 -fPIC -ffp-contract=off -w -shared \
    -I"$HERE/cvshim" -I"$REF/include" \
    "$REF/src/ORBextractor.cc" "$HERE/ref_driver.cpp" "$HERE/cvprims.cpp" \
    -o "$HERE/_ref/libref_orbextractor.so"
echo "built $HERE/_ref/libref_orbextractor.so"
