/*
 * oracle/port/orc_fft.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement of the float FFT / MDCT filterbank slots of FFTContext (libavcodec/fft.h:73-99):
 *   fft_permute + fft_calc  libavcodec/fft_template.c:186-194, :140-346  == natural-order DFT, forward kernel
 *                           exp(-2 pi i jk / n), inverse exp(+...), no 1/n (reference test: libavcodec/tests/fft.c:72-118)
 *   imdct_half / imdct_calc / mdct_calc   libavcodec/mdct_template.c:95-214 with the pre/post rotation tables of
 *                           ff_mdct_init (:86-92): theta = 1/8 (+ n/4 if scale < 0), tcos/tsin = -cos/-sin(alpha) * sqrt(|scale|)
 * The transform core here is an iterative radix-2 FFT in DOUBLE precision (results rounded to float once), not the
 * reference's float split-radix: the contract for this row is a tolerance (north star: 1e-6 relative), and a double
 * oracle sits between the reference's float result and the GPU's float result.  Pinned against oracle/_ref in
 * tests/test_oracle_fft_cpu.py at that tolerance.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include "../oracle_api.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct { double re, im; } cd;

/* natural order in, natural order out; sign = -1 forward, +1 inverse */
static void dft_pow2(cd *z, int nbits, int sign)
{
    int n = 1 << nbits;
    for (int i = 0, j = 0; i < n; i++) {                       /* bit reversal */
        if (i < j) { cd t = z[i]; z[i] = z[j]; z[j] = t; }
        int m = n >> 1;
        while (m >= 1 && (j & m)) { j ^= m; m >>= 1; }
        j |= m;
    }
    for (int len = 2; len <= n; len <<= 1) {
        double ang = sign * 2.0 * M_PI / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; k++) {
                double c = cos(ang * k), s = sin(ang * k);
                cd a = z[i + k], b = z[i + k + len / 2];
                cd t = { b.re * c - b.im * s, b.re * s + b.im * c };
                z[i + k].re = a.re + t.re; z[i + k].im = a.im + t.im;
                z[i + k + len / 2].re = a.re - t.re; z[i + k + len / 2].im = a.im - t.im;
            }
    }
}

void orc_fft(int nbits, int inverse, float *zf)
{
    int n = 1 << nbits;
    cd *z = malloc(sizeof(cd) * n);
    for (int i = 0; i < n; i++) { z[i].re = zf[2 * i]; z[i].im = zf[2 * i + 1]; }
    dft_pow2(z, nbits, inverse ? 1 : -1);
    for (int i = 0; i < n; i++) { zf[2 * i] = (float)z[i].re; zf[2 * i + 1] = (float)z[i].im; }
    free(z);
}

/* rotation tables exactly as ff_mdct_init builds them (double math, stored as float) */
static void mdct_tables(int nbits, double scale, float *tcos, float *tsin)
{
    int n = 1 << nbits, n4 = n >> 2;
    double theta = 1.0 / 8.0 + (scale < 0 ? n4 : 0), sc = sqrt(fabs(scale));
    for (int i = 0; i < n4; i++) {
        double alpha = 2 * M_PI * (i + theta) / n;
        tcos[i] = (float)(-cos(alpha) * sc);
        tsin[i] = (float)(-sin(alpha) * sc);
    }
}
#define CMULD(dre, dim, are, aim, bre, bim) do { (dre) = (double)(are) * (bre) - (double)(aim) * (bim); (dim) = (double)(are) * (bim) + (double)(aim) * (bre); } while (0)

void orc_imdct_half(int nbits, double scale, float *out, const float *in)
{
    int n = 1 << nbits, n2 = n >> 1, n4 = n >> 2, n8 = n >> 3;
    float *tcos = malloc(sizeof(float) * n2), *tsin = tcos + n4;
    cd *z = malloc(sizeof(cd) * n4);
    mdct_tables(nbits, scale, tcos, tsin);
    for (int k = 0; k < n4; k++) CMULD(z[k].re, z[k].im, in[n2 - 1 - 2 * k], in[2 * k], tcos[k], tsin[k]);   /* pre rotation */
    dft_pow2(z, nbits - 2, 1);
    for (int k = 0; k < n8; k++) {                                                                          /* post rotation */
        double r0, i0, r1, i1;
        CMULD(r0, i1, z[n8 - k - 1].im, z[n8 - k - 1].re, tsin[n8 - k - 1], tcos[n8 - k - 1]);
        CMULD(r1, i0, z[n8 + k].im, z[n8 + k].re, tsin[n8 + k], tcos[n8 + k]);
        out[2 * (n8 - k - 1)] = (float)r0; out[2 * (n8 - k - 1) + 1] = (float)i0;
        out[2 * (n8 + k)] = (float)r1; out[2 * (n8 + k) + 1] = (float)i1;
    }
    free(z); free(tcos);
}

void orc_imdct_calc(int nbits, double scale, float *out, const float *in)
{
    int n = 1 << nbits, n2 = n >> 1, n4 = n >> 2;
    orc_imdct_half(nbits, scale, out + n4, in);
    for (int k = 0; k < n4; k++) { out[k] = -out[n2 - k - 1]; out[n - k - 1] = out[n2 + k]; }
}

void orc_mdct_calc(int nbits, double scale, float *out, const float *in)
{
    int n = 1 << nbits, n2 = n >> 1, n4 = n >> 2, n8 = n >> 3, n3 = 3 * n4;
    float *tcos = malloc(sizeof(float) * n2), *tsin = tcos + n4;
    cd *x = malloc(sizeof(cd) * n4);
    mdct_tables(nbits, scale, tcos, tsin);
    for (int i = 0; i < n8; i++) {
        double re = -in[2 * i + n3] - in[n3 - 1 - 2 * i], im = -in[n4 + 2 * i] + in[n4 - 1 - 2 * i];
        CMULD(x[i].re, x[i].im, re, im, -tcos[i], tsin[i]);
        re = in[2 * i] - in[n2 - 1 - 2 * i]; im = -in[n2 + 2 * i] - in[n - 1 - 2 * i];
        CMULD(x[n8 + i].re, x[n8 + i].im, re, im, -tcos[n8 + i], tsin[n8 + i]);
    }
    dft_pow2(x, nbits - 2, -1);
    for (int i = 0; i < n8; i++) {
        double r0, i0, r1, i1;
        CMULD(i1, r0, x[n8 - i - 1].re, x[n8 - i - 1].im, -tsin[n8 - i - 1], -tcos[n8 - i - 1]);
        CMULD(i0, r1, x[n8 + i].re, x[n8 + i].im, -tsin[n8 + i], -tcos[n8 + i]);
        out[2 * (n8 - i - 1)] = (float)r0; out[2 * (n8 - i - 1) + 1] = (float)i0;
        out[2 * (n8 + i)] = (float)r1; out[2 * (n8 + i) + 1] = (float)i1;
    }
    free(x); free(tcos);
}
