#!/bin/bash
# Builds the reference's ONLY native component (src/libs/cutils.pyx, Cython) from the sources
# where they lie under /root/reference, into oracle/_ref/cutils.so (git-ignored).
#
# The reference is Python + this one Cython module.  Intermediates (the patched .pyx and the
# generated .c) live in a temp dir and are deleted; only the compiled .so lands in oracle/_ref/.
# One token differs from upstream: Cython 3's numpy.pxd renamed NPY_OWNDATA -> NPY_ARRAY_OWNDATA
# (cutils.pyx:22); without it the module raises AttributeError at the first call.
#
# The .so is used (a) by tests/golden/make_golden.py to run the reference's own Python layers in
# this container and (b) by tests to validate oracle/cutils_port.c against the real thing.
# It is never imported by the product path.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/src/libs/cutils.pyx" ]; then
  echo "build_ref: $REF not present (GPU box?) - keeping whatever is in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
sed 's/np\.NPY_OWNDATA/np.NPY_ARRAY_OWNDATA/' "$REF/src/libs/cutils.pyx" > "$TMP/cutils.pyx"
PY=${PYTHON:-python}
"$PY" -m cython -3 "$TMP/cutils.pyx" -o "$TMP/cutils.c"
INC_PY="$("$PY" -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
INC_NP="$("$PY" -c 'import numpy; print(numpy.get_include())')"
gcc -O2 -fPIC -shared -fopenmp -w -DNPY_NO_DEPRECATED_API=0 -I"$INC_PY" -I"$INC_NP" "$TMP/cutils.c" -o "$OUT/cutils.so"
echo "build_ref: wrote $OUT/cutils.so"
