#!/bin/bash
# usage: sweep_variants.sh name1 name2 ...   ("default" = the in-tree library)
for v in "$@"; do
  if [ "$v" = default ]; then unset MFB_LIB_PATH; else export MFB_LIB_PATH=$PWD/tools/scratch/variants/libmfb_$v.so; fi
  echo "== $v"; python tools/time_bwd.py 2>&1 | grep tape=
done
