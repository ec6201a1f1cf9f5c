#!/usr/bin/env bash
# Build the UNMODIFIED reference (libmorpho 0.6.4) from the sources where they
# lie under /root/reference/src into oracle/_ref/ (git-ignored, but shipped to
# the GPU box by gpurun).  Recipe: SURVEY.md Appendix A.  The reference's own
# CMake build is NOT used (it needs system cblas/lapacke/SuiteSparse headers
# which are absent); instead every src/*/*.c except the dead xmatrix.c is
# compiled directly with three prototype-only shim headers (shim/), a
# from-scratch CSparse subset (shim/cs_stub.c) and the OpenBLAS 0.3.15 that
# ships in this image inside the opencv wheel.
#
# Outputs (all under oracle/_ref/):
#   libmorpho_ref.so        the reference library
#   morpho6                 our 60-line CLI driver (morpho6_main.c)
#   libmorpho_refharness.so our ctypes-callable harness (ref_harness.c)
#   pkg/share/modules/*.morpho   copy of the reference's script modules so that
#                                `import optimize` works on the GPU box
#   home/.morphopackages    package list pointing at pkg/ (resources.c:246-273)
#   reftests/                copy of the reference's test/{functionals,selection,field,mesh} (run on the GPU box too)
# TEST INFRASTRUCTURE ONLY.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${MORPHO_REFERENCE:-/root/reference}"
SRC="$REF/src"
OUT="$REPO/oracle/_ref"
OBJ="$OUT/obj"
# absolute path that is valid here and on the GPU box (/root/repo is a symlink there)
PKG="/root/repo/oracle/_ref/pkg"

if [ ! -d "$SRC" ]; then
  echo "build_ref.sh: $SRC not present (GPU box?) — using prebuilt oracle/_ref" >&2
  exit 0
fi
OB="$(python3 - <<'PY'
import glob, os, sys
import importlib.util
cands = []
for p in sys.path:
    cands += glob.glob(os.path.join(p, "opencv_python_headless.libs", "libopenblasp-*.so"))
print(cands[0] if cands else "")
PY
)"
if [ -z "$OB" ]; then echo "build_ref.sh: OpenBLAS not found in the image" >&2; exit 1; fi
OBDIR="$(dirname "$OB")"; OBLIB="$(basename "$OB")"

mkdir -p "$OBJ" "$OUT/pkg/share/modules" "$OUT/pkg/lib" "$OUT/pkg/share/help" "$OUT/home"
INC="-I$HERE/shim -I$SRC"
for d in builtin classes core datastructures debug geometry linalg support; do INC="$INC -I$SRC/$d"; done
DEFS="-DMORPHO_INCLUDE_LINALG -DMORPHO_INCLUDE_SPARSE -DMORPHO_INCLUDE_GEOMETRY -DMORPHO_LINALG_USE_CSPARSE"
DEFS="$DEFS -DMORPHO_DYLIBEXTENSION=\"so\" -DMORPHO_MODULE_BASEDIR=\"$PKG/share/modules\" -DMORPHO_HELP_BASEDIR=\"$PKG/share/help\""
CFLAGS="-O2 -fPIC -std=gnu11 -w"

build_one() {
  local f="$1" o="$OBJ/$(echo "${1#$SRC/}" | tr '/' '_' | sed 's/\.c$/.o/')"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then gcc $CFLAGS $INC $DEFS -c "$f" -o "$o"; fi
}
export -f build_one; export SRC OBJ CFLAGS INC DEFS
find "$SRC" -name '*.c' ! -path '*/linalg/xmatrix.c' | sort | xargs -P "$(nproc)" -I{} bash -c 'build_one {}'
gcc $CFLAGS -I"$HERE/shim" -c "$HERE/shim/cs_stub.c" -o "$OBJ/cs_stub.o"

gcc -shared -o "$OUT/libmorpho_ref.so" "$OBJ"/*.o -L"$OBDIR" -l:"$OBLIB" -Wl,--disable-new-dtags -Wl,-rpath,"$OBDIR" -lm -ldl -lpthread
gcc $CFLAGS $INC $DEFS "$HERE/morpho6_main.c" -o "$OUT/morpho6" -L"$OUT" -lmorpho_ref \
    -Wl,--disable-new-dtags -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$OBDIR" -Wl,-rpath-link,"$OBDIR" -lm
gcc $CFLAGS $INC $DEFS -shared "$HERE/ref_harness.c" -o "$OUT/libmorpho_refharness.so" -L"$OUT" -lmorpho_ref \
    -Wl,--disable-new-dtags -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$OBDIR" -Wl,-rpath-link,"$OBDIR" -lm

cp -f "$REF"/modules/*.morpho "$OUT/pkg/share/modules/"
# the reference's own golden-output tests of the hot path, so that the GPU box (which has no /root/reference) can run
# them through the extension as well (tests/test_extension.py).  Build output only: oracle/_ref is never committed.
rm -rf "$OUT/reftests"; mkdir -p "$OUT/reftests"
for d in functionals selection field mesh; do cp -r "$REF/test/$d" "$OUT/reftests/$d"; done
# the shipped examples that exercise the hot path (BASELINE.json configs: catenoid, thomson, tactoid; and qtensor, cholesteric,
# electrostatics, cube), run AS SHIPPED by the tests
rm -rf "$OUT/examples"; mkdir -p "$OUT/examples"
for d in catenoid thomson tactoid qtensor cholesteric electrostatics cube; do cp -r "$REF/examples/$d" "$OUT/examples/$d"; done
rm -f "$OUT"/examples/qtensor/*.pdf "$OUT"/examples/qtensor/src.*
echo "$PKG" > "$OUT/home/.morphopackages"
echo "$OBDIR" > "$OUT/openblas_dir.txt"
echo "build_ref.sh: built $(ls "$OBJ"/*.o | wc -l) objects -> $OUT"
