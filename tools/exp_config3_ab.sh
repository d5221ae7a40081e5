#!/bin/bash
# A/B of prebuilt library variants on config 3 (mode 9 + 40 fields + augmentation): tools/exp_config3_ab.sh variants/a.so ...
LIB=optical-flow-2d-data-generation_b200/csrc/libofdg.so
cp $LIB /tmp/libofdg_current.so
for v in "$@"; do
  cp "$v" $LIB
  echo "[$v]" $(python tools/exp_config3.py 50 1 2>&1 | tail -1)
  echo "[$v]" $(python tools/exp_config3.py 50 0 2>&1 | tail -1)
done
cp /tmp/libofdg_current.so $LIB
