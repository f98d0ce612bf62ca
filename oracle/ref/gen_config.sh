#!/bin/sh
# Writes a hand-rolled config.h / config_components.h / avconfig.h / ffversion.h for a
# generic-C (no arch asm) build of the few reference source files the DSP hot paths need.
# This is NOT the reference's configure: every HAVE_/CONFIG_/ARCH_ symbol that the selected
# sources mention is defined to 0, then the short list below is switched to 1.
# usage: gen_config.sh <reference-root> <out-include-dir> <file-list...>
set -e
REF="$1"; OUT="$2"; shift 2
mkdir -p "$OUT/libavutil"
ONES="HAVE_THREADS HAVE_PTHREADS HAVE_FAST_UNALIGNED HAVE_FAST_64BIT HAVE_FAST_CLZ HAVE_LOCAL_ALIGNED
HAVE_ATTRIBUTE_PACKED HAVE_ATTRIBUTE_MAY_ALIAS HAVE_PRAGMA_DEPRECATED HAVE_BUILTIN_VECTOR
HAVE_UNISTD_H HAVE_SYS_TIME_H HAVE_GETTIMEOFDAY HAVE_CLOCK_GETTIME HAVE_NANOSLEEP HAVE_USLEEP
HAVE_POSIX_MEMALIGN HAVE_MEMALIGN HAVE_ALIGNED_MALLOC_DISABLED HAVE_SYSCONF HAVE_SCHED_GETAFFINITY HAVE_ISATTY
HAVE_LRINT HAVE_LRINTF HAVE_RINT HAVE_RINTF HAVE_ROUND HAVE_ROUNDF HAVE_TRUNC HAVE_TRUNCF HAVE_CBRT HAVE_CBRTF
HAVE_COPYSIGN HAVE_ERF HAVE_EXP2 HAVE_EXP2F HAVE_EXPF HAVE_HYPOT HAVE_ISFINITE HAVE_ISINF HAVE_ISNAN
HAVE_LDEXPF HAVE_LLRINT HAVE_LLRINTF HAVE_LOG2 HAVE_LOG2F HAVE_LOG10F HAVE_POWF HAVE_SINF HAVE_COSF HAVE_ATANF
HAVE_ATAN2F HAVE_LDEXPF HAVE_SYMVER HAVE_INLINE_ASM_DISABLED HAVE_STRUCT_TIMESPEC_DISABLED
HAVE_IO_H_DISABLED HAVE_MKSTEMP HAVE_LOCALTIME_R HAVE_GMTIME_R HAVE_STRERROR_R HAVE_ACCESS HAVE_FCNTL HAVE_LSTAT
HAVE_SYS_RESOURCE_H HAVE_GETRUSAGE HAVE_MMAP HAVE_MPROTECT HAVE_ERRNO_H_PLACEHOLDER HAVE_BIGENDIAN_DISABLED
HAVE_GETENV HAVE_SECURE_GETENV_DISABLED HAVE_STDBIT_H_DISABLED
CONFIG_SWSCALE CONFIG_AVUTIL CONFIG_AVCODEC CONFIG_SMALL_DISABLED CONFIG_MEMORY_POISONING_DISABLED
CONFIG_IDCTDSP CONFIG_ME_CMP CONFIG_H264QPEL CONFIG_HPELDSP CONFIG_FAANIDCT CONFIG_SWSCALE_ALPHA CONFIG_SAFE_BITSTREAM_READER CONFIG_PIXELUTILS
CONFIG_GPL CONFIG_UNSTABLE_DISABLED"
SYMS=$(cd "$REF" && cat "$@" libavutil/*.h libswscale/*.h libavcodec/idctdsp.h libavcodec/me_cmp.h compat/*.h 2>/dev/null \
       | grep -ohE '\b(HAVE|CONFIG|ARCH)_[A-Za-z0-9_]+\b' | sort -u)
# symbols only reachable through token pasting (HAVE_ ## ext ## suffix in libavutil/cpu_internal.h)
for e in MMX MMXEXT SSE SSE2 SSE3 SSSE3 SSE4 SSE42 AVX AVX2 AVX512 AVX512ICL FMA3 FMA4 XOP AESNI AMD3DNOW AMD3DNOWEXT \
         NEON ARMV8 VFP VFPV3 ARMV5TE ARMV6 ARMV6T2 SETEND DOTPROD I8MM SVE SVE2 SME ALTIVEC VSX POWER8 LSX LASX MSA MMI RVV RV \
         RV_ZVBB RV_MISALIGNED; do
  SYMS="$SYMS HAVE_$e HAVE_${e}_EXTERNAL HAVE_${e}_INLINE"
done
SYMS=$(echo $SYMS | tr ' ' '\n' | sort -u)
{
  echo "/* hand-rolled by oracle/ref/gen_config.sh — generic C build, no arch asm */"
  echo "#ifndef FFREF_CONFIG_H"; echo "#define FFREF_CONFIG_H"
  echo '#define FFMPEG_CONFIGURATION "b200-oracle generic-c"'
  echo '#define FFMPEG_LICENSE "LGPL version 2.1 or later"'
  echo '#define CONFIG_THIS_YEAR 2026'
  echo '#define FFMPEG_DATADIR "/nonexistent"'
  echo '#define AVCONV_DATADIR "/nonexistent"'
  echo '#define CC_IDENT "gcc"'
  echo '#define OS_NAME linux'
  echo '#define EXTERN_PREFIX ""'
  echo '#define EXTERN_ASM '
  echo '#define BUILDSUF ""'
  echo '#define SLIBSUF ".so"'
  echo '#define SWS_MAX_FILTER_SIZE 256'
  echo '#define av_restrict restrict'
  for s in $SYMS; do
    v=0
    for o in $ONES; do [ "$o" = "$s" ] && v=1; done
    echo "#define $s $v"
  done
  echo "#endif"
} > "$OUT/config.h"
# components: nothing enabled (no codecs/filters are built)
{
  echo "#ifndef FFREF_CONFIG_COMPONENTS_H"; echo "#define FFREF_CONFIG_COMPONENTS_H"
  (cd "$REF" && cat "$@" | grep -ohE '\bCONFIG_[A-Z0-9_]+_(DECODER|ENCODER|PARSER|FILTER|MUXER|DEMUXER|HWACCEL|BSF|PROTOCOL|INDEV|OUTDEV)\b' | sort -u) \
    | while read s; do echo "#ifndef $s"; echo "#define $s 0"; echo "#endif"; done
  echo "#endif"
} > "$OUT/config_components.h"
cat > "$OUT/libavutil/avconfig.h" <<EOT
#ifndef AVUTIL_AVCONFIG_H
#define AVUTIL_AVCONFIG_H
#define AV_HAVE_BIGENDIAN 0
#define AV_HAVE_FAST_UNALIGNED 1
#endif
EOT
cat > "$OUT/libavutil/ffversion.h" <<EOT
#ifndef AVUTIL_FFVERSION_H
#define AVUTIL_FFVERSION_H
#define FFMPEG_VERSION "b200-oracle"
#endif
EOT
