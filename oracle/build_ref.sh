#!/usr/bin/env bash
# TEST INFRASTRUCTURE: builds the reference's own CPU path (NeuralNetwork<Cpu> + layers::*<Cpu>)
# from the sources where they lie under /root/reference into oracle/_ref/libcurrennt_ref.so.
#
# No reference source is copied into the repository: a scratch copy is made under $TMPDIR,
# patched mechanically for the CUDA 12.9 toolchain, compiled, and deleted.  Only the shared
# library lands in oracle/_ref/ (git-ignored; it still travels to the GPU box).
#
# Mechanical fixes:
#   1. Thrust in CUDA 12.9 aliases thrust::tuple to cuda::std::tuple, which has no member
#      get<N>() -> rewrite `x.get<N>()` to `thrust::get<N>(x)` in the scratch copy.
#   2. thrust::transform_reduce needs its own header -> -include thrust/transform_reduce.h
#   3. Boost is absent -> oracle/boost_shim maps the 7 headers used onto the C++ standard library.
#   4. WeightedSsePostOutputLayer / SseMaskPostOutputLayer::calculateError reduce over bare counting iterators, which Thrust
#      dispatches to its DEVICE backend even in the Cpu instantiation (host pointers on the GPU: it cannot run as written
#      under --cuda false).  Those two translation units are compiled with THRUST_DEVICE_SYSTEM=CPP so that the same source
#      runs its sequential host reduction; nothing else in them depends on the backend.
#   5. Configuration.cpp / data_sets/DataSet.cpp need Boost.program_options / libnetcdf and are
#      not on the hot path -> their few referenced symbols live in oracle/ref_harness/ref_harness.cu.
set -euo pipefail

HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${CURRENNT_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
# REF_VARIANT=omp builds the same sources with Thrust's OpenMP host backend (all host cores; timing baseline only -- its parallel
# reductions sum in a different order, so parity always uses the default sequential build)
VARIANT="${REF_VARIANT:-}"
LIB="$OUT/libcurrennt_ref${VARIANT:+_$VARIANT}.so"

if [ ! -d "$REF/currennt_lib/src" ]; then
    if [ -f "$LIB" ]; then echo "[build_ref] $REF absent; keeping prebuilt $LIB"; exit 0; fi
    echo "[build_ref] $REF absent and no prebuilt library" >&2; exit 3
fi

# up to date?
if [ -f "$LIB" ] && [ "$LIB" -nt "$HERE/ref_harness/ref_harness.cu" ] && [ "$LIB" -nt "$HERE/build_ref.sh" ]; then
    echo "[build_ref] $LIB up to date"; exit 0
fi

NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
WORK="$(mktemp -d "${TMPDIR:-/tmp}/currennt_ref_build.XXXXXX")"
trap 'rm -rf "$WORK"' EXIT
mkdir -p "$OUT" "$WORK/src" "$WORK/obj"

cp -r "$REF/currennt_lib/src/." "$WORK/src/"
chmod -R u+w "$WORK/src"
# fix 1: tuple.get<N>() -> thrust::get<N>(tuple)
find "$WORK/src/layers" "$WORK/src/optimizers" "$WORK/src/helpers" -name '*.cu' -print0 |
    xargs -0 sed -E -i 's/\b([A-Za-z_]+)\.get<([0-9])>\(\)/thrust::get<\2>(\1)/g'

SRCS=(
    layers/Layer.cpp layers/InputLayer.cpp layers/PostOutputLayer.cpp
    layers/TrainableLayer.cu layers/FeedForwardLayer.cu layers/SoftmaxLayer.cu layers/LstmLayer.cu
    layers/SsePostOutputLayer.cu layers/RmsePostOutputLayer.cu layers/CePostOutputLayer.cu
    layers/SseMaskPostOutputLayer.cu layers/WeightedSsePostOutputLayer.cu
    layers/BinaryClassificationLayer.cu layers/MulticlassClassificationLayer.cu
    helpers/Matrix.cu helpers/cublas.cu helpers/JsonClasses.cpp
    data_sets/DataSetFraction.cpp NeuralNetwork.cpp LayerFactory.cu
)

FLAGS=(-x cu -O3 -Xcompiler -O3,-fPIC -std=c++17 -arch=sm_100a -w
       -include thrust/transform_reduce.h -I "$HERE/boost_shim" -I "$WORK/src")
LINK=()
if [ "$VARIANT" = "omp" ]; then
    FLAGS+=(-DTHRUST_HOST_SYSTEM=THRUST_HOST_SYSTEM_OMP -Xcompiler -fopenmp)
    LINK=(-Xcompiler -fopenmp)
fi

pids=()
for s in "${SRCS[@]}"; do
    o="$WORK/obj/$(echo "$s" | tr '/.' '__').o"
    extra=()
    case "$s" in layers/SseMaskPostOutputLayer.cu|layers/WeightedSsePostOutputLayer.cu)
        extra=(-DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP) ;; esac
    "$NVCC" "${FLAGS[@]}" "${extra[@]}" -c "$WORK/src/$s" -o "$o" &
    pids+=($!)
    # at most $(nproc) compilers at once
    while [ "$(jobs -rp | wc -l)" -ge "$(nproc)" ]; do sleep 0.2; done
done
"$NVCC" "${FLAGS[@]}" -c "$HERE/ref_harness/ref_harness.cu" -o "$WORK/obj/ref_harness.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done

"$NVCC" -shared -Xlinker -Bsymbolic -arch=sm_100a "${LINK[@]}" -o "$LIB" "$WORK"/obj/*.o -lcublas
echo "[build_ref] built $LIB"
