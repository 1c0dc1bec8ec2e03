#!/usr/bin/env python3
"""Write a hand-rolled config.h / avconfig.h for compiling the *unmodified* libav
sources where they lie (read-only) into oracle/_ref/.  TEST INFRASTRUCTURE ONLY.

We do not run the reference's configure.  Instead every ARCH_/HAVE_/CONFIG_ token that
the reference sources mention is defined to 0, except a short allow-list that describes
this host (little-endian x86-64 Linux, gcc, pthreads, libm).  ARCH_X86 stays 0 on
purpose: the oracle is the reference's portable C path, which is the parity target
named by the north star (idct_algo=FF_IDCT_SIMPLE, SWS_BITEXACT|SWS_ACCURATE_RND).
"""
import os, re, sys

REF = sys.argv[1]
OUT = sys.argv[2]
SIMD = len(sys.argv) > 3 and sys.argv[3] == "simd"     # the timing-only x86 flavour (oracle/_ref/libavref_simd.so), see Makefile

ONES = set("""
HAVE_FAST_64BIT HAVE_FAST_CLZ HAVE_FAST_CMOV HAVE_FAST_UNALIGNED
HAVE_LOCAL_ALIGNED_8 HAVE_LOCAL_ALIGNED_16 HAVE_LOCAL_ALIGNED_32
HAVE_SIMD_ALIGN_16 HAVE_ALIGNED_STACK
HAVE_ATANF HAVE_ATAN2F HAVE_CBRTF HAVE_COSF HAVE_EXP2 HAVE_EXP2F HAVE_EXPF HAVE_ISINF
HAVE_ISNAN HAVE_LDEXPF HAVE_LLRINT HAVE_LLRINTF HAVE_LOG2 HAVE_LOG2F HAVE_LOG10F
HAVE_LRINT HAVE_LRINTF HAVE_POWF HAVE_RINT HAVE_ROUND HAVE_ROUNDF HAVE_SINF HAVE_TRUNC
HAVE_TRUNCF HAVE_THREADS HAVE_PTHREADS HAVE_CLOCK_GETTIME HAVE_GETTIMEOFDAY
HAVE_GMTIME_R HAVE_LOCALTIME_R HAVE_POSIX_MEMALIGN HAVE_MEMALIGN HAVE_MALLOC_H
HAVE_UNISTD_H HAVE_SYS_TIME_H HAVE_SYS_PARAM_H HAVE_SYSCONF HAVE_USLEEP HAVE_NANOSLEEP
HAVE_SCHED_GETAFFINITY HAVE_STRERROR_R HAVE_FCNTL HAVE_MMAP HAVE_ISATTY
HAVE_PRAGMA_DEPRECATED HAVE_GETOPT HAVE_SYNC_VAL_COMPARE_AND_SWAP HAVE_ATTRIBUTE_PACKED
HAVE_ATTRIBUTE_MAY_ALIAS
CONFIG_GPL CONFIG_FAANDCT CONFIG_FAANIDCT CONFIG_FDCTDSP CONFIG_IDCTDSP CONFIG_BLOCKDSP
CONFIG_ME_CMP CONFIG_H264DSP CONFIG_H264QPEL CONFIG_H264CHROMA CONFIG_HPELDSP
CONFIG_FFT CONFIG_MDCT CONFIG_RDFT CONFIG_DCT CONFIG_SWSCALE CONFIG_AVUTIL CONFIG_AVCODEC
CONFIG_SMALL_NOT
""".split())
ONES.discard("CONFIG_SMALL_NOT")
if SIMD:
    # what the reference's configure finds on an x86-64 gcc host when no external assembler is present (--disable-x86asm): the
    # inline-asm MMX / MMXEXT / SSE / SSE2 / SSSE3 code is compiled in and selected at run time from cpuid; every *_EXTERNAL stays 0
    ONES |= set("""
    ARCH_X86 ARCH_X86_64 HAVE_INLINE_ASM HAVE_MMX HAVE_MMXEXT HAVE_SSE HAVE_SSE2 HAVE_SSE3 HAVE_SSSE3
    HAVE_MMX_INLINE HAVE_MMXEXT_INLINE HAVE_SSE_INLINE HAVE_SSE2_INLINE HAVE_SSE3_INLINE HAVE_SSSE3_INLINE
    HAVE_XMM_CLOBBERS HAVE_EBP_AVAILABLE HAVE_EBX_AVAILABLE HAVE_INLINE_ASM_LABELS HAVE_INLINE_ASM_NONLOCAL_LABELS
    HAVE_INLINE_ASM_DIRECT_SYMBOL_REFS HAVE_I686 HAVE_RDTSC
    """.split())

tok = re.compile(r"\b((?:ARCH|HAVE|CONFIG)_[A-Z0-9_]+)\b")
names = set()
for sub in ("libavutil", "libavcodec", "libswscale", "compat"):
    for root, _, files in os.walk(os.path.join(REF, sub)):
        for f in files:
            if f.endswith((".c", ".h")):
                with open(os.path.join(root, f), errors="replace") as fh:
                    names.update(tok.findall(fh.read()))
names |= ONES
# the x86 cpu-flag macros are token-pasted (libavutil/x86/cpu.h, cpu_internal.h): name every (extension, flavour) pair explicitly
for ext in "MMX MMXEXT SSE SSE2 SSE3 SSSE3 SSE4 SSE42 AVX AVX2 XOP FMA3 FMA4 AMD3DNOW AMD3DNOWEXT".split():
    names |= {"HAVE_" + ext, "HAVE_%s_EXTERNAL" % ext, "HAVE_%s_INLINE" % ext}
os.makedirs(os.path.join(OUT, "libavutil"), exist_ok=True)
with open(os.path.join(OUT, "config.h"), "w") as o:
    o.write("/* written by oracle/refbuild/gen_config.py - not the reference's configure */\n")
    o.write("#ifndef LIBAV_CONFIG_H\n#define LIBAV_CONFIG_H\n")
    o.write('#define LIBAV_CONFIGURATION "oracle-refbuild"\n#define LIBAV_LICENSE "GPL version 2 or later"\n')
    o.write('#define CC_IDENT "gcc"\n#define EXTERN_PREFIX ""\n#define EXTERN_ASM\n#define SLIBSUF ".so"\n')
    o.write('#define restrict __restrict__\n') if False else None
    for n in sorted(names):
        o.write("#define %s %d\n" % (n, 1 if n in ONES else 0))
    o.write("#endif\n")
with open(os.path.join(OUT, "libavutil", "avconfig.h"), "w") as o:
    o.write("#ifndef AVUTIL_AVCONFIG_H\n#define AVUTIL_AVCONFIG_H\n"
            "#define AV_HAVE_BIGENDIAN 0\n#define AV_HAVE_FAST_UNALIGNED 1\n#endif\n")
with open(os.path.join(OUT, "avversion.h"), "w") as o:
    o.write('#define LIBAV_VERSION "oracle-refbuild"\n')
print("config.h: %d macros, %d set" % (len(names), len(ONES & names)))
