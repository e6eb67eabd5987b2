#!/bin/bash
# tools/var_sweep.sh -- device-resident rate of library variants built into tools/var_*.so (A/B of compile-time options)
L=orb-slam3_byzyh_b200/libORBfe_b200.so
cp $L /tmp/lib_default.so
for v in default "$@" default; do
  if [ $v = default ]; then cp /tmp/lib_default.so $L; else cp tools/var_$v.so $L; fi
  echo "== $v"; python tools/resident_rate.py "" | tail -1
done
cp /tmp/lib_default.so $L
