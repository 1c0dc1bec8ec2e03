#!/bin/sh
# Builds the reference's own FFT/MDCT self test (libavcodec/tests/fft.c, FATE targets fate-fft-N, fate-ifft-N,
# fate-mdct-N, fate-imdct-N for N = 4..12, tests/fate/fft.mak) against oracle/_ref/libavref.so, runs every target and
# records the verdicts in tests/golden/fate_fft.txt (the program itself applies FATE's 1e-3 threshold and exits
# non-zero when it is exceeded).  Run in the container that has /root/reference; the output file is committed.
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd); REF=${REF:-/root/reference}
OUT=$ROOT/oracle/_ref
make -s -C "$HERE" REF="$REF"
gcc -O2 -std=c99 -D_ISOC99_SOURCE -D_POSIX_C_SOURCE=200112 -D_XOPEN_SOURCE=600 -D_DEFAULT_SOURCE -DHAVE_AV_CONFIG_H \
    -I"$OUT/cfg" -I"$REF" -w -o "$OUT/fft_selftest" "$REF/libavcodec/tests/fft.c" "$REF/libavcodec/rdft.c" "$REF/libavcodec/dct.c" "$REF/libavcodec/dct32_float.c" -L"$OUT" -lavref -lm -Wl,-rpath,"$OUT"
G=$ROOT/tests/golden/fate_fft.txt
echo "# libavcodec/tests/fft (float build) against oracle/_ref: FATE target, exit status (0 = within FATE's 1e-3 threshold)" > "$G"
for n in 4 5 6 7 8 9 10 11 12; do
    for t in "fft:" "ifft:-i" "mdct:-m" "imdct:-m -i"; do
        name=${t%%:*}; args=${t#*:}
        if "$OUT/fft_selftest" -n$n $args > /dev/null 2>&1; then rc=0; else rc=$?; fi
        echo "fate-$name-$n $rc" >> "$G"
    done
done
cat "$G" | head -12
