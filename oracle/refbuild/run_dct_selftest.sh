#!/bin/sh
# Builds the reference's own IDCT self test (libavcodec/tests/dct.c, FATE target fate-idct8x8) against
# oracle/_ref/libavref.so and records its SIMPLE-C lines as known answers in tests/golden/fate_idct8x8.txt.
# Run in the container that has /root/reference; the output file is committed.
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd); REF=${REF:-/root/reference}
OUT=$ROOT/oracle/_ref
make -s -C "$HERE" REF="$REF"
gcc -O2 -std=c99 -D_ISOC99_SOURCE -D_POSIX_C_SOURCE=200112 -D_XOPEN_SOURCE=600 -DHAVE_AV_CONFIG_H \
    -I"$OUT/cfg" -I"$REF" -w -o "$OUT/dct_selftest" "$REF/libavcodec/tests/dct.c" "$REF/libavcodec/dctref.c" \
    "$REF/libavcodec/aandcttab.c" "$REF/libavcodec/xvididct.c" -L"$OUT" -lavref -lm -Wl,-rpath,"$OUT"
G=$ROOT/tests/golden/fate_idct8x8.txt
echo "# libavcodec/tests/dct -i <test>: the 'IDCT SIMPLE-C' line (NB_ITS=20000, av_lfg seed 1)" > "$G"
for t in 0 1 2; do
    "$OUT/dct_selftest" -i $t | grep 'IDCT SIMPLE-C' | sed "s/^IDCT SIMPLE-C:/test$t/" >> "$G"
done
cat "$G"
