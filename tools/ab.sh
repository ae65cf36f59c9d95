#!/bin/bash
# A/B of library variants built into torchregister_b200/_lib/variants/*.so: tools/ab.sh <script.py> [args]
set -e
L=torchregister_b200/_lib
cp $L/libtrb_b200.so /tmp/_orig.so
for v in $L/variants/*.so; do
  n=$(basename $v .so)
  cp $v $L/libtrb_b200.so
  timeout 300 python "$@" $n 2>&1 | tail -8
done
cp /tmp/_orig.so $L/libtrb_b200.so
