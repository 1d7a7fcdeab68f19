#!/bin/bash
# Relink the reference's own tools against this repo's library (for real):
#   tools/_build/astcenc-b200      the reference command line tool (Source/astcenccli_*.cpp, compiled where they lie, with
#                                  include/astcenc.h of THIS repo force-included instead of Source/astcenc.h) linked against
#                                  astc-encoder_b200/libastcenc_b200.so
#   tools/_build/astcenc-ref       the same tool linked against the unmodified reference library (oracle/_ref) - the
#                                  comparison partner of the GPU test
#   tools/_build/unittests-b200    Source/UnitTest/test_encode.cpp + test_decode.cpp (GoogleTest from Source/GoogleTest),
#                                  compiled against include/astcenc.h and linked against libastcenc_b200.so
# Nothing of the reference is copied: sources are compiled from /root/reference, outputs go to tools/_build (git-ignored,
# travels to the GPU box). The reference's cmake build is not run; the link rule mirrors Source/cmake_core.cmake:97-110
# (tool = cli sources + veneers + the library; here the library is ours).
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF=${REF:-/root/reference}
OUT="$REPO/tools/_build"
SRC="$REF/Source"
if [ ! -d "$SRC" ]; then echo "reference sources absent: keeping prebuilt $OUT"; exit 0; fi
mkdir -p "$OUT/obj" "$OUT/gen"
cat > "$OUT/gen/astcenccli_version.h" <<'H'
#ifndef ASTCENCCLI_VERSION_INCLUDED
#define ASTCENCCLI_VERSION_INCLUDED
#define VERSION_STRING "5.6.0-b200-relink"
#define YEAR_STRING "2026"
#endif
H
CXX=${CXX:-g++}
BASE="-std=c++14 -O2 -DNDEBUG -ffp-contract=off -Wno-deprecated-declarations -DASTCENC_NEON=0 -DASTCENC_SVE=0 -DASTCENC_SSE=41 -DASTCENC_AVX=2 -DASTCENC_POPCNT=1 -DASTCENC_F16C=1 -mavx2 -mpopcnt -mf16c -I$OUT/gen -I$SRC"
# our public header replaces the reference's: force-include it and pre-define the reference header's include guard
OURS="-include $REPO/include/astcenc.h -DASTCENC_INCLUDED"
CLI="astcenccli_entry astcenccli_entry2 astcenccli_error_metrics astcenccli_image astcenccli_image_external astcenccli_image_load_store astcenccli_platform_dependents astcenccli_toplevel astcenccli_toplevel_help"
# library internals the tool borrows (half-float conversion, maths helpers): two reference files, compiled into the tool
LIBBITS="astcenc_mathlib astcenc_mathlib_softfloat"
pids=""
for f in $CLI $LIBBITS; do
  ( $CXX $BASE $OURS -c "$SRC/$f.cpp" -o "$OUT/obj/b200_$f.o" ) & pids="$pids $!"
  ( $CXX $BASE -c "$SRC/$f.cpp" -o "$OUT/obj/ref_$f.o" ) & pids="$pids $!"
done
GT="$SRC/GoogleTest/googletest"
( $CXX -std=c++14 -O1 -I"$GT/include" -I"$GT" -c "$GT/src/gtest-all.cc" -o "$OUT/obj/gtest-all.o" ) & pids="$pids $!"
( $CXX -std=c++14 -O1 -I"$GT/include" -I"$GT" -c "$GT/src/gtest_main.cc" -o "$OUT/obj/gtest_main.o" ) & pids="$pids $!"
for t in test_encode test_decode; do
  ( $CXX -std=c++14 -O1 -I"$GT/include" $OURS -c "$SRC/UnitTest/$t.cpp" -o "$OUT/obj/ut_$t.o" ) & pids="$pids $!"
done
for p in $pids; do wait $p; done
B200_OBJS=""; REF_OBJS=""
for f in $CLI $LIBBITS; do B200_OBJS="$B200_OBJS $OUT/obj/b200_$f.o"; REF_OBJS="$REF_OBJS $OUT/obj/ref_$f.o"; done
$CXX -o "$OUT/astcenc-b200" $B200_OBJS -L"$REPO/astc-encoder_b200" -lastcenc_b200 -Wl,-rpath,'$ORIGIN/../../astc-encoder_b200' -lpthread
$CXX -o "$OUT/unittests-b200" "$OUT/obj/ut_test_encode.o" "$OUT/obj/ut_test_decode.o" "$OUT/obj/gtest-all.o" "$OUT/obj/gtest_main.o" \
     -L"$REPO/astc-encoder_b200" -lastcenc_b200 -Wl,-rpath,'$ORIGIN/../../astc-encoder_b200' -lpthread
if [ -f "$REPO/oracle/_ref/libastcenc_ref_avx2.so" ]; then
  # (the reference library was built with hidden visibility + ASTCENC_DYNAMIC_LIBRARY: it exports the same ten symbols)
  $CXX -o "$OUT/astcenc-ref" $REF_OBJS "$REPO/oracle/_ref/libastcenc_ref_avx2.so" -Wl,-rpath,'$ORIGIN/../../oracle/_ref' -lpthread
fi
# no astcenc_* symbol may stay unresolved outside the two libraries
for exe in astcenc-b200 unittests-b200; do
  und=$(nm -D --undefined-only "$OUT/$exe" | grep -E " astcenc_" | awk '{print $2}' | sort -u)
  for s in $und; do
    nm -D --defined-only "$REPO/astc-encoder_b200/libastcenc_b200.so" | grep -q " $s$" || { echo "UNRESOLVED in $exe: $s"; exit 1; }
  done
  echo "$exe: $(echo $und | wc -w) astcenc_* imports, all provided by libastcenc_b200.so"
done
ls -la "$OUT" | grep -E "astcenc-|unittests"
