#!/bin/sh
# TEST INFRASTRUCTURE. Builds oracle/_ref/libofdg_ref.so: the REFERENCE'S OWN sources, compiled untouched from
# where they lie under $REF (default /root/reference), + oracle/ref_api.cpp (a C interface for the tests).
#
#   sh oracle/ref_build.sh                        # third-party arithmetic from oracle/shim (restated AGG 2.4 / CImg)
#   AGG_INCLUDE=/path/to/agg-2.4/include AGG_LIB=/path/to/agg-2.4/src/libagg.a \
#   CIMG_INCLUDE=/path/holding/thirdparty/CImg/CImg.h  sh oracle/ref_build.sh
#                                                 # the real libraries, when someone can supply the two downloads the
#                                                 # reference asks for (cmake/Dependencies.cmake:4-22, README.md:38);
#                                                 # real include directories are searched BEFORE the shim's
#
# Caffe (LMB fork), protobuf, glog and boost are always shimmed (oracle/shim/caffe, oracle/shim/boost): the generator does
# no arithmetic through them. Flags: the reference builds as part of Caffe's Release configuration (-O2 -DNDEBUG; the
# assert at data_generation_layer.cpp:127 does not even parse otherwise); -ffp-contract=off and no -march so float/double
# expressions round like the plain x86-64 SSE build the reference is; CPU_ONLY because its Forward_gpu is Forward_cpu.
# Outputs go to oracle/_ref/ only (git-ignored; travels to the GPU box with the snapshot).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${REF:-/root/reference}
OUT=$HERE/_ref
CXX=${REF_CXX:-g++}   # not $CXX: some images point it at a wrapper that links libstdc++ statically, which breaks iostreams inside a dlopen()ed library
[ -f "$REF/src/caffe/DataGenerator.cpp" ] || { echo "ref_build: no reference sources under $REF" >&2; exit 3; }
mkdir -p "$OUT"
INC=""
[ -n "$AGG_INCLUDE" ] && INC="$INC -I$AGG_INCLUDE"
[ -n "$CIMG_INCLUDE" ] && INC="$INC -I$CIMG_INCLUDE"
INC="$INC -I$HERE/shim -I$HERE/shim/agg -I$REF/include -I$HERE -I$HERE/../include"
FLAGS="-std=c++14 -O2 -fPIC -ffp-contract=off -DNDEBUG -DCPU_ONLY -w -pthread"
for f in src/caffe/DataGenerator.cpp src/caffe/WarpFields.cpp src/caffe/layers/data_generation_layer.cpp; do
  o=$OUT/$(basename "$f" .cpp).o
  if [ ! -f "$o" ] || [ "$REF/$f" -nt "$o" ] || [ "$HERE/shim/agg/agg_shim.h" -nt "$o" ] || [ "$HERE/shim/thirdparty/CImg/CImg.h" -nt "$o" ] \
     || [ "$HERE/shim/caffe/ofdg_caffe_shim.hpp" -nt "$o" ] || [ "$HERE/shim/caffe/proto/caffe.pb.h" -nt "$o" ] || [ "$0" -nt "$o" ]; then
    $CXX $FLAGS $INC -c "$REF/$f" -o "$o" &
  fi
done
wait
$CXX $FLAGS -fno-access-control $INC -c "$HERE/ref_api.cpp" -o "$OUT/ref_api.o"
$CXX -shared -pthread -Wl,-Bsymbolic-functions -o "$OUT/libofdg_ref.so.tmp" "$OUT/DataGenerator.o" "$OUT/WarpFields.o" "$OUT/data_generation_layer.o" "$OUT/ref_api.o" $AGG_LIB
mv "$OUT/libofdg_ref.so.tmp" "$OUT/libofdg_ref.so"
echo "$OUT/libofdg_ref.so"
