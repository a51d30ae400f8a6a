#!/usr/bin/env bash
# Builds libliesel_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
       -I ../../include -Xptxas -v --expt-relaxed-constexpr ${NVCC_EXTRA:-})
SRCS=(plan.cu tmap.cu tiny.cu logp_grad.cu iwls.cu group_iwls.cu spline_basis.cu debug.cu banded.cu banded_x2.cu dense_x2.cu banded_iwls_x2.cu dgroup_x2.cu msmooth_x2.cu msort_x2.cu nuts.cu nuts_async.cu dense_tc.cu peer.cu)
# tools library (pipe-peak micro-benchmarks, tcgen05 probe): NOT linked into the product library
TOOL_SRCS=(microbench.cu tc_probe.cu)
mkdir -p build
pids=()
for s in "${SRCS[@]}" "${TOOL_SRCS[@]}"; do
  o=build/${s%.cu}.o
  if [[ ! -f $o || $s -nt $o || common.cuh -nt $o || prologue.cuh -nt $o || x2_common.cuh -nt $o || tc_bf16.cuh -nt $o || ../../include/liesel_b200.h -nt $o ]]; then
    ( $NVCC "${FLAGS[@]}" -c "$s" -o "$o" > "build/${s%.cu}.log" 2>&1 || { cat "build/${s%.cu}.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
PROD_OBJS=(); for s in "${SRCS[@]}"; do PROD_OBJS+=("build/${s%.cu}.o"); done
TOOL_OBJS=(); for s in "${TOOL_SRCS[@]}"; do TOOL_OBJS+=("build/${s%.cu}.o"); done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o libliesel_b200.so "${PROD_OBJS[@]}" -lcudart
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o libliesel_b200_bench.so "${TOOL_OBJS[@]}" -L. -lliesel_b200 -lcudart -Xlinker -rpath -Xlinker '$ORIGIN'
echo "built $(pwd)/libliesel_b200_bench.so"
echo "built $(pwd)/libliesel_b200.so"
# The jax.ffi shim, compiled against the TEST STUB of xla/ffi/api/ffi.h (jaxlib is not installable
# here): keeps the shim compiling and lets tests/test_ffi_shim.py drive its handlers.
g++ -shared -fPIC -std=c++17 -O1 -Wall -I ../../tests/stubs -I ../../include -I /usr/local/cuda/include \
    xla_ffi_shim.cc -L. -lliesel_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/..' \
    -o build/liblslb_xla_ffi_stub.so
echo "built $(pwd)/build/liblslb_xla_ffi_stub.so"
