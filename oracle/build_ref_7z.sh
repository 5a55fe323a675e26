#!/bin/bash
# oracle/build_ref_7z.sh -- TEST INFRASTRUCTURE ONLY.
# Builds the UNMODIFIED reference 7-Zip console host and codec/format module from /root/reference (copied to a scratch
# directory because its makefiles write into the tree), in two flavours, into oracle/_ref/7z/ (git-ignored, travels to the
# GPU box):
#   stock/7zz            the stand-alone reference (Bundles/Alone2): verifies what the GPU codecs write, makes test archives
#   host/7z, host/7z.so  UI/Console + Bundles/Format7zF with ONLY the registrations of ZSTD (4F71101), LZMA2 (21) and FLZMA2 (21)
#                        left out of the object list (Arc_gcc.mak:261,417,418 -- ZstdRegister.o, Lzma2Register.o,
#                        FastLzma2Register.o), so that the host looks these methods up in Codecs/*.so -- where the tests put
#                        libb200z_7z.so (CPP/7zip/UI/Common/LoadCodecs.cpp:528-563, Common/CreateCoder.cpp:159-232: internal
#                        codecs win over external ones, which is why a stock 7z.so cannot be overridden side by side).
# No reference source is copied into the repo; nothing here is linked into the product.
set -e
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref/7z"
[ -d "$REF/CPP/7zip" ] || { echo "build_ref_7z: $REF absent -- using prebuilt $OUT if present"; exit 0; }
if [ -x "$OUT/stock/7zz" ] && [ -x "$OUT/host/7z" ] && [ -f "$OUT/host/7z.so" ] && [ "$1" != "-f" ]; then exit 0; fi
W=$(mktemp -d /tmp/b2z_ref7z.XXXXXX)
trap 'rm -rf "$W"' EXIT
cp -r "$REF/C" "$REF/CPP" "$REF/Asm" "$W/"
J=${J:-$(nproc)}
mkdir -p "$OUT/stock" "$OUT/host/Codecs"
( cd "$W/CPP/7zip/Bundles/Alone2" && make -j"$J" -f makefile.gcc FLAGS_FLTO= > "$W/alone2.log" 2>&1 ) || { tail -30 "$W/alone2.log"; exit 1; }
cp "$W/CPP/7zip/Bundles/Alone2/_o/7zz" "$OUT/stock/7zz"
# the codec-less module: drop the three registrations from the object list (the coder objects stay; nothing registers them)
sed -i -e '/\$O\/Lzma2Register\.o/d' -e '/\$O\/ZstdRegister\.o/d' -e '/\$O\/FastLzma2Register\.o/d' "$W/CPP/7zip/Bundles/Format7zF/Arc_gcc.mak"
( cd "$W/CPP/7zip/Bundles/Format7zF" && make -j"$J" -f makefile.gcc FLAGS_FLTO= > "$W/format7zf.log" 2>&1 ) || { tail -30 "$W/format7zf.log"; exit 1; }
cp "$W/CPP/7zip/Bundles/Format7zF/_o/7z.so" "$OUT/host/7z.so"
( cd "$W/CPP/7zip/UI/Console" && make -j"$J" -f makefile.gcc FLAGS_FLTO= > "$W/console.log" 2>&1 ) || { tail -30 "$W/console.log"; exit 1; }
cp "$W/CPP/7zip/UI/Console/_o/7z" "$OUT/host/7z"
echo "built $OUT/stock/7zz, $OUT/host/7z + 7z.so (ZSTD / LZMA2 / FLZMA2 registrations left out)"
