#!/bin/bash
# Compiles the REFERENCE's own alias-table path (ZetaCore/Math/Sampling.cpp + Math/Common.cpp and the
# headers they pull in) where it lies under /root/reference into oracle/_ref/libref_alias.so.
#
# No reference source is copied into the repo. Two LLP64-only literals (`4llu`, `1llu`; size_t is
# unsigned long on LP64) stop gcc, so the two affected files are patched in a scratch copy under
# $TMPDIR which is deleted afterwards; the MSVC-only _aligned_malloc/_aligned_free and __forceinline /
# __vectorcall keywords are shimmed from the command line. NDEBUG: the reference's debug Assert macro
# does not compile under gcc (Utility/Error.h:40).
set -euo pipefail
REF=${ZR_REFERENCE_DIR:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/../_ref"
if [ ! -d "$REF/Source/ZetaCore" ]; then
    echo "reference tree not found at $REF -- keeping any prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/ZetaCore"
cp -r "$REF/Source/ZetaCore/Math" "$REF/Source/ZetaCore/Utility" "$REF/Source/ZetaCore/Support" \
      "$REF/Source/ZetaCore/App" "$REF/Source/ZetaCore/Win32" "$TMP/ZetaCore/" 2>/dev/null || true
chmod -R u+w "$TMP"
sed -i 's/4llu/(size_t)4/g' "$TMP/ZetaCore/Utility/SmallVector.h"
sed -i 's/1llu/(size_t)1/g' "$TMP/ZetaCore/Math/Common.cpp"
cat > "$TMP/shim.h" <<'EOF'
#pragma once
#include <cstdlib>
#include <cstddef>
static inline void* _aligned_malloc(size_t size, size_t alignment)
{
    void* p = nullptr;
    if (alignment < sizeof(void*)) alignment = sizeof(void*);
    if (posix_memalign(&p, alignment, size) != 0) return nullptr;
    return p;
}
static inline void _aligned_free(void* p) { free(p); }
EOF
g++ -std=c++20 -O2 -fPIC -shared -mavx2 -mfma -mf16c -ffp-contract=off -DNDEBUG \
    -D__forceinline=inline -D__vectorcall= -fpermissive -w \
    -include "$TMP/shim.h" -I"$TMP/ZetaCore" -I"$TMP" -I"$REF/External" \
    "$TMP/ZetaCore/Math/Sampling.cpp" "$TMP/ZetaCore/Math/Common.cpp" "$HERE/ref_alias_wrap.cpp" \
    -o "$OUT/libref_alias.so"
echo "built $OUT/libref_alias.so"
