#!/bin/bash
# round 2, call R: approxmatch unroll factors (sweeps / final pass)
V=$PWD/monopsr_b200/build/variants
echo "base (sweep unroll 8, final 1)"; timeout 100 python tools/am_quick.py 2>&1 | grep "scale 1.0"
for u in u16f1 u8f2 u8f4 u16f2; do echo $u; MPB_LIB=$V/lib_$u.so timeout 100 python tools/am_quick.py 2>&1 | grep "scale 1.0"; done
