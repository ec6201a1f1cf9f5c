#!/usr/bin/env bash
# Build the Morpho extension morphocuda.so (morpho_b200/ext/morphocuda.c) against the reference's headers and
# libmorpho_cuda.so.  Needs /root/reference (headers) and oracle/_ref (the reference library the test driver
# morpho6 links): run by __graft_entry__.build() in the build container; the GPU box uses the prebuilt file.
# Outputs: morpho_b200/lib/morphocuda.so and a copy in the test package root oracle/_ref/pkg/lib/.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${MORPHO_REFERENCE:-/root/reference}"
SRC="$REF/src"
if [ ! -d "$SRC" ]; then echo "build_ext.sh: $SRC not present — using the prebuilt morphocuda.so" >&2; exit 0; fi
INC="-I$REPO/oracle/refbuild/shim -I$SRC -I$REPO/include"
for d in builtin classes core datastructures debug geometry linalg support; do INC="$INC -I$SRC/$d"; done
DEFS="-DMORPHO_INCLUDE_LINALG -DMORPHO_INCLUDE_SPARSE -DMORPHO_INCLUDE_GEOMETRY -DMORPHO_LINALG_USE_CSPARSE"
LIBDIR="$REPO/morpho_b200/lib"
mkdir -p "$LIBDIR" "$REPO/oracle/_ref/pkg/lib"
# rpath: the repository path is the same here and on the GPU box (/root/repo is a symlink there)
gcc -O2 -fPIC -std=gnu11 -Wall -Wno-unused-function $INC $DEFS -shared "$HERE/morphocuda.c" -o "$LIBDIR/morphocuda.so" \
    -L"$LIBDIR" -lmorpho_cuda -Wl,--disable-new-dtags -Wl,-rpath,/root/repo/morpho_b200/lib -Wl,-rpath,'$ORIGIN'
cp -f "$LIBDIR/morphocuda.so" "$REPO/oracle/_ref/pkg/lib/morphocuda.so"
echo "build_ext.sh: built $LIBDIR/morphocuda.so"
