/* oracle/refcfg/julius/config.h -- TEST INFRASTRUCTURE (oracle build only).
 *
 * Engine switch list for compiling the UNMODIFIED libjulius sources where
 * they lie under /root/reference (see oracle/Makefile).  Hand-written: the
 * stock "fast" setup (UNIGRAM_FACTORING, LOWMEM2,
 * PASS1_IWCD, GPRUNE_DEFAULT_BEAM, SCAN_BEAM, CONFIDENCE_MEASURE,
 * LM_FIX_DOUBLE_SCORING, ENABLE_PLUGIN; not WPAIR / WORD_GRAPH / DETERMINE).
 */
#ifndef JB200_ORACLE_JULIUS_CONFIG_H
#define JB200_ORACLE_JULIUS_CONFIG_H
#define JULIUS_PRODUCTNAME "JuliusLib"
#define JULIUS_VERSION "4.6"
#define JULIUS_SETUP "fast"
#define JULIUS_HOSTINFO "x86_64-unknown-linux-gnu"
#define JULIUS_BUILD_INFO "gcc -O6 -fomit-frame-pointer -fPIC (oracle/Makefile)"
#define RETSIGTYPE void
#define STDC_HEADERS 1
#define HAVE_PTHREAD 1
#define UNIGRAM_FACTORING 1
#define LOWMEM2 1
#define PASS1_IWCD 1
#define SCAN_BEAM 1
#define GPRUNE_DEFAULT_BEAM 1
#define CONFIDENCE_MEASURE 1
#define LM_FIX_DOUBLE_SCORING 1
#define GRAPHOUT_DYNAMIC 1
#define GRAPHOUT_SEARCH 1
#define ENABLE_PLUGIN 1
#define HAVE_LIBFVAD 1
#endif
