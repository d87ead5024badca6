#!/bin/bash
# A/B kernel variants on one box: bash scripts/ab.sh "<name>:<nvcc -D flags>" ...   (build here, run under gpurun)
# build: bash scripts/ab.sh build name:flags ... ; run: bash scripts/ab.sh run name ...
MODE=$1; shift
mkdir -p build/variants
if [ "$MODE" = build ]; then
  for v in "$@"; do
    name=${v%%:*}; flags=${v#*:}
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
      -o build/variants/lib_$name.so alphazero-rs_b200/csrc/engine.cu || exit 1
  done
else
  for rep in 1 2 3; do
    for name in "$@"; do
      for g in 592 4096; do
        echo -n "$name games=$g: "
        AZB200_LIB=build/variants/lib_$name.so python scripts/profile_selfplay.py $g | sed 's/.*device_ms.: \([0-9.]*\).*/\1 ms/'
      done
    done
  done
fi
