/* oracle/refcfg/sent/config.h -- TEST INFRASTRUCTURE (oracle build only).
 *
 * Build-time switch list for compiling the UNMODIFIED libsent sources where
 * they lie under /root/reference (see oracle/Makefile).  Written by hand for
 * this repo: it states the switches of Julius' stock "fast" setup on
 * linux/x86-64 (the values its configure script would pick with
 * --with-mictype=oss), because the reference's own build system is not run.
 * Audio capture is irrelevant to the hot path; only file / vector input is
 * exercised by the oracle.
 */
#ifndef JB200_ORACLE_SENT_CONFIG_H
#define JB200_ORACLE_SENT_CONFIG_H
#define LIBSENT_VERSION "4.6"
#define AUDIO_API_NAME "oss"
#define AUDIO_API_DESC "Open Sound System compatible"
#define AUDIO_FORMAT_DESC "RAW and WAV only"
#define GZIP_READING_DESC "zlib library"
#define STDC_HEADERS 1
#define USE_MIC 1
#define USE_ADDLOG_ARRAY 1
#define HAVE_SOCKLEN_T 1
#define HAVE_UNISTD_H 1
#define HAVE_ZLIB 1
#define HAVE_STRCASECMP 1
#define HAVE_SLEEP 1
#define CLASS_NGRAM 1
#define MFCC_SINCOS_TABLE 1
#define HAVE_SYS_SOUNDCARD_H 1
#define HAS_OSS 1
#define USE_MBR 1
#define HAS_SIMD_FMA 1
#define HAS_SIMD_AVX 1
#define HAS_SIMD_SSE 1
#endif
