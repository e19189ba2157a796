/* jb200_dl.h -- run-time binding of libjb200.so from C host code (plugin / beam shim).
 * The library is located through $JB200_LIB, else next to this object:  <dir>/../../julius_b200/libjb200.so */
#ifndef JB200_DL_H
#define JB200_DL_H
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "julius_b200.h"

typedef struct {
  void *so;
  const char *(*last_error)(void);
  int (*gmm_create)(const jb200_gmm_desc *, int, int, jb200_gmm **);
  int (*gmm_score_host)(jb200_gmm *, const float *, int, float *);
  int (*gmm_gauss_host)(jb200_gmm *, const float *, float *);
  int (*dnn_create)(const jb200_dnn_desc *, int, jb200_dnn **);
  int (*dnn_score_host)(jb200_dnn *, const float *, int, float *);
  int (*decoder_create)(const jb200_tree_desc *, jb200_gmm *, int, int, jb200_decoder **);
  int (*decoder_attach_dnn)(jb200_decoder *, jb200_dnn *);
  int (*decode_batch_host)(jb200_decoder *, const float *, const int32_t *, int);
  int (*decoder_results)(jb200_decoder *, const jb200_utt_result **, const jb200_atom **, const int32_t **);
  void (*decoder_destroy)(jb200_decoder *);
  int (*stream_open)(jb200_decoder *, int);
  int (*stream_feed_host)(jb200_decoder *, const float *, const int32_t *, const uint8_t *, int);
  int (*stream_status)(jb200_decoder *, int, int32_t *, int32_t *, int32_t *);
  int (*stream_partial)(jb200_decoder *, int, int32_t *, int, int32_t *, float *, int32_t *);
  int (*stream_result)(jb200_decoder *, int, const jb200_utt_result **, const jb200_atom **, const int32_t **);
  void (*gmm_destroy)(jb200_gmm *);
  void (*dnn_destroy)(jb200_dnn *);
} jb200_api;

static int jb200_api_load(jb200_api *a, void *anchor) {
  char path[4096];
  const char *env = getenv("JB200_LIB");
  Dl_info info;
  if (env) snprintf(path, sizeof(path), "%s", env);
  else if (dladdr(anchor, &info) && info.dli_fname) {
    char *slash;
    snprintf(path, sizeof(path), "%s", info.dli_fname);
    slash = strrchr(path, '/');
    if (slash) *slash = '\0'; else strcpy(path, ".");
    strncat(path, "/../../julius_b200/libjb200.so", sizeof(path) - strlen(path) - 1);
  } else snprintf(path, sizeof(path), "libjb200.so");
  a->so = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!a->so) { fprintf(stderr, "jb200: cannot load %s: %s\n", path, dlerror()); return -1; }
#define JB200_SYM(field, name) do { *(void **)(&a->field) = dlsym(a->so, name); if (!a->field) { fprintf(stderr, "jb200: %s lacks %s\n", path, name); return -1; } } while (0)
  JB200_SYM(last_error, "jb200_last_error");
  JB200_SYM(gmm_create, "jb200_gmm_create");
  JB200_SYM(gmm_score_host, "jb200_gmm_score_host");
  JB200_SYM(gmm_gauss_host, "jb200_gmm_gauss_host");
  JB200_SYM(dnn_create, "jb200_dnn_create");
  JB200_SYM(dnn_score_host, "jb200_dnn_score_host");
  JB200_SYM(decoder_create, "jb200_decoder_create");
  JB200_SYM(decoder_attach_dnn, "jb200_decoder_attach_dnn");
  JB200_SYM(decode_batch_host, "jb200_decode_batch_host");
  JB200_SYM(decoder_results, "jb200_decoder_results");
  JB200_SYM(decoder_destroy, "jb200_decoder_destroy");
  JB200_SYM(stream_open, "jb200_stream_open");
  JB200_SYM(stream_feed_host, "jb200_stream_feed_host");
  JB200_SYM(stream_status, "jb200_stream_status");
  JB200_SYM(stream_partial, "jb200_stream_partial");
  JB200_SYM(stream_result, "jb200_stream_result");
  JB200_SYM(gmm_destroy, "jb200_gmm_destroy");
  JB200_SYM(dnn_destroy, "jb200_dnn_destroy");
#undef JB200_SYM
  return 0;
}
#endif
