#!/bin/bash
for v in "" mb3 mb5; do
  if [ -z "$v" ]; then unset MFB_LIB_PATH; echo "== default (4 CTAs/SM, 128 regs)"; else export MFB_LIB_PATH=$PWD/monoforce_b200/libmfb_$v.so; echo "== $v"; fi
  python tools/quick_time.py 2>&1 | grep -E "fwd forces|fwd no-forces"
done
