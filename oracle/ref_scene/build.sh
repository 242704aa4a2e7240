#!/bin/bash
# Compiles the REFERENCE's own scene-data headers (Core/Material.h, RayTracing/RtCommon.h, Math/*) where they lie under
# /root/reference, behind oracle/ref_scene/ref_scene_wrap.cpp, into oracle/_ref/libref_scene.so. Same recipe as
# oracle/ref_alias/build.sh (scratch copy for the one LLP64 literal, MSVC keywords shimmed on the command line, NDEBUG).
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
mkdir -p "$TMP/ZetaCore" "$TMP/ZetaRenderPass"
for d in Math Utility Support App Win32 Core RayTracing; do cp -r "$REF/Source/ZetaCore/$d" "$TMP/ZetaCore/" 2>/dev/null || true; done
cp -r "$REF/Source/ZetaRenderPass/Common" "$TMP/ZetaRenderPass/"
chmod -R u+w "$TMP"
sed -i 's/4llu/(size_t)4/g' "$TMP/ZetaCore/Utility/SmallVector.h"
cat > "$TMP/shim.h" <<'EOS'
#pragma once
#include <cstdlib>
#include <cstddef>
#include <cstring>
static inline void* _aligned_malloc(size_t size, size_t alignment)
{
    void* p = nullptr;
    if (alignment < sizeof(void*)) alignment = sizeof(void*);
    if (posix_memalign(&p, alignment, size) != 0) return nullptr;
    return p;
}
static inline void _aligned_free(void* p) { free(p); }
EOS
g++ -std=c++20 -O2 -fPIC -shared -mavx2 -mfma -mf16c -ffp-contract=off -DNDEBUG \
    -D__forceinline=inline -D__vectorcall= -fpermissive -w \
    -include "$TMP/shim.h" -I"$TMP/ZetaCore" -I"$TMP" -I"$REF/External" \
    "$HERE/ref_scene_wrap.cpp" -o "$OUT/libref_scene.so"
echo "built $OUT/libref_scene.so"
