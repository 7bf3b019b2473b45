#!/usr/bin/env bash
# Build recipe for the UNMODIFIED reference (CommediaJW/Dist-GNN) as a differential
# oracle for the GPU parity tests.  TEST INFRASTRUCTURE ONLY - nothing under
# dist-gnn_b200/ may import or link what this produces.
#
# The reference sources are compiled where they lie under $REF (default
# /root/reference, read-only); nothing is copied into this repo.  Objects go to a
# scratch dir, the only output in-tree is oracle/_ref/dgs.cpython-*.so (git-ignored,
# travels to the GPU box with the gpurun snapshot).  The reference's own cmake build
# is NOT run: this is a flat nvcc compile of the file list in its CMakeLists.txt:44-56.
#
# Toolchain-drift handling (no kernel logic is touched, no source is edited):
#   * -std=c++17            (torch 2.11 headers need C++17; reference asks for 14)
#   * -include thrust/...   (CCCL no longer pulls these in transitively)
#   * c10::Storage::data() now returns const void*: the three host files that assign it to a
#     void* (src/common/pin_memory.cc:9,17, src/nccl/nccl_context.cc:95,98,
#     src/cache/tensor_p2p_cache.cc:42) are compiled from a sed-patched scratch copy in $OBJ
#     (outside the repo) with "storage().data()" -> "storage().mutable_data()"; their relative
#     includes still resolve into $REF through -I.
#   * -DNDEBUG is NOT set   (device asserts stay live, as in the reference's build)
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
OBJ=${OBJ:-/tmp/dgs_ref_obj}
PY=${PYTHON:-python}
if [ ! -d "$REF/src" ]; then echo "build_ref: $REF not present - skipping (prebuilt $OUT is used if it exists)"; exit 0; fi
mkdir -p "$OUT" "$OBJ"
TORCH_DIR=$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))' 2>/dev/null | tail -1)
PYINC=$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')
PBINC=$($PY -c 'import pybind11;print(pybind11.get_include())')
SUFFIX=$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')
NCCL_LIBDIR=$($PY -c 'import nvidia.nccl,os;print(os.path.join(list(nvidia.nccl.__path__)[0],"lib"))' 2>/dev/null || true)
SRCS=$(cd "$REF" && ls src/*.cc src/cache/*.cc src/cache/cuda/*.cu src/context/*.cc src/common/*.cc \
        src/hashmap/cuda/*.cu src/nccl/*.cc src/sampling/*.cc src/sampling/cuda/*.cu \
        src/feature/*.cc src/feature/cuda/*.cu)
FLAGS="-x cu -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a --expt-extended-lambda --expt-relaxed-constexpr \
 -Xcompiler -fPIC,-w -w \
 -include thrust/execution_policy.h -include thrust/for_each.h -include thrust/iterator/counting_iterator.h \
 -DTORCH_EXTENSION_NAME=dgs -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1 \
 -I$REF/include -I$TORCH_DIR/include -I$TORCH_DIR/include/torch/csrc/api/include -I$PYINC -I$PBINC -I/usr/include"
pids=()
objs=()
for s in $SRCS; do
  o=$OBJ/$(echo "$s" | tr '/' '_').o
  objs+=("$o")
  src="$REF/$s"
  extra=""
  case "$s" in
    src/common/pin_memory.cc|src/nccl/nccl_context.cc|src/cache/tensor_p2p_cache.cc)
      mkdir -p "$OBJ/patched/$(dirname "$s")"
      sed 's/storage()\.data()/storage().mutable_data()/g' "$REF/$s" > "$OBJ/patched/$s"
      src="$OBJ/patched/$s"; extra="-I$REF/$(dirname "$s")";;
  esac
  if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ]; then
    ( nvcc $FLAGS $extra -c "$src" -o "$o" 2> "$o.log" || { echo "FAILED $s"; tail -30 "$o.log"; exit 1; } ) &
    pids+=($!)
    # bound parallelism
    while [ "$(jobs -rp | wc -l)" -ge "${JOBS:-8}" ]; do sleep 0.5; done
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
nvcc -shared -o "$OUT/dgs$SUFFIX" "${objs[@]}" \
  -L"$TORCH_DIR/lib" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -ltorch_python \
  ${NCCL_LIBDIR:+-L$NCCL_LIBDIR} -l:libnccl.so.2 -lcudart \
  -Xlinker -rpath,"$TORCH_DIR/lib" ${NCCL_LIBDIR:+-Xlinker -rpath,$NCCL_LIBDIR}
echo "build_ref: wrote $OUT/dgs$SUFFIX"
