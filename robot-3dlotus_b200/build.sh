#!/bin/bash
# Build liblotus.so (all sm_100a kernels + the C ABI) in-tree.  Usage: robot-3dlotus_b200/build.sh [extra nvcc flags]
# Every .cu compiles to its own object in parallel (objects are rebuilt only when the source or a header is newer), then one link.
set -e
cd "$(dirname "$0")/csrc"
mkdir -p ../_obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden $*"
echo "$FLAGS" > ../_obj/.flags.new
if ! cmp -s ../_obj/.flags.new ../_obj/.flags 2>/dev/null; then rm -f ../_obj/*.o; mv ../_obj/.flags.new ../_obj/.flags; fi
newest_hdr=$(ls -t *.cuh ../../include/*.h 2>/dev/null | head -1)
pids=()
for f in *.cu; do
  o=../_obj/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$newest_hdr" -a "$newest_hdr" -nt "$o" ]; then
    nvcc $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC,-fvisibility=hidden -o ../liblotus.so ../_obj/*.o
