/* jb200_attach.c -- in-process attach of the GPU acoustic scorer to a running Julius engine.
 *
 * Part of the .jpi plugin (built with jb200_export.c).  With JB200_ATTACH=1 the startup() hook
 *   1. flattens the live models (jb200_export.c),
 *   2. loads libjb200.so and creates the GPU scorer (GMM, exact arithmetic by default; or DNN),
 *   3. registers a CALLBACK_EVENT_PASS1_BEGIN handler (libjulius/include/julius/callback.h:119) that,
 *      for buffered input (all T frames present), scores [T x S] on the GPU in one call and stores the
 *      rows in HMMWork.outprob_cache (libsent/include/sent/hmm_calc.h:115).  Every later
 *      outprob_state() (libsent/src/phmm/outprob.c:183-249) is a cache hit: pass 1, pass 2 and
 *      -outprobout all consume the GPU's numbers.  This is the "HMMWork function pointers / cache"
 *      boundary of SURVEY.md 8b.
 * It also provides the calcmix hook set so that "-gprune jb200" is a valid jconf value
 * (libjulius/src/plugin.c:336-354, contract plugin/calcmix.c:226-323): per-Gaussian ln scores of the
 * current frame come from the GPU (jb200_gmm_gauss_host), one device call per frame, sliced per state.
 * JB200_ATTACH=calcmix attaches only this hook (no cache fill), JB200_ATTACH=1 the whole-utterance scoring.
 * JB200_GMM_MODE=fast selects the FMA/exact-LSE arithmetic (<=1e-4) instead of the bit-exact one.
 */
#include <julius/juliuslib.h>
#include "jb200_model.h"
#include "jb200_dl.h"

static jb200_api g_api;
static jb200_gmm *g_gmm = NULL;
static jb200_dnn *g_dnn = NULL;
static jb200_gmm_desc g_gd;
static jb200_dnn_desc g_dd;
static float *g_scores = NULL; static size_t g_scores_cap = 0;
/* calcmix state */
static float *g_gauss = NULL; static int g_gauss_time = -2; static const int32_t *g_state_off = NULL;

static void on_pass1_begin(Recog *recog, void *dummy) {
  PROCESS_AM *am = recog->amlist;
  HMMWork *wrk = &(am->hmmwrk);
  HTK_Param *param = am->mfcc->param;
  int T = param->samplenum, S = wrk->statenum, D, t, rc;
  float *in;
  if (T <= 0 || param->is_outprob) return;
  if (recog->jconf->decodeopt.realtime_flag) {
    jlog("WARNING: jb200: real-time (frame-by-frame) input: the per-utterance GPU scoring is skipped\n");
    return;
  }
  D = g_dnn ? g_dd.in_dim : g_gd.dim;
  if (param->veclen < D) { jlog("ERROR: jb200: parameter vector shorter than the model's input\n"); return; }
  if ((size_t)T * S > g_scores_cap) { g_scores_cap = (size_t)T * S; g_scores = (float *)realloc(g_scores, sizeof(float) * g_scores_cap); }
  /* parvec rows are separate allocations in general: gather into one matrix */
  in = (float *)malloc(sizeof(float) * (size_t)T * D);
  for (t = 0; t < T; t++) memcpy(in + (size_t)t * D, param->parvec[t], sizeof(float) * D);
  rc = g_dnn ? g_api.dnn_score_host(g_dnn, in, T, g_scores) : g_api.gmm_score_host(g_gmm, in, T, g_scores);
  free(in);
  if (rc != 0) { jlog("ERROR: jb200: GPU scoring failed: %s\n", g_api.last_error()); return; }
  /* make the cache rows exist (outprob_cache_extend is static: outprob.c:116), then overwrite them */
  outprob_state(wrk, T - 1, am->hmminfo->ststart, param);
  for (t = 0; t < T; t++) memcpy(wrk->outprob_cache[t], g_scores + (size_t)t * S, sizeof(float) * S);
  wrk->OP_time = -1;          /* force outprob_state() to re-latch its per-frame pointers */
  wrk->OP_last_time = -1;
}

/* calcmix-only attach: a new utterance must not reuse the last utterance's frame of Gaussian scores */
static void on_pass1_begin_calcmix(Recog *recog, void *dummy) { g_gauss_time = -2; }

int jb200_attach(Recog *recog, jb200_blob *b) {
  const char *mode = getenv("JB200_GMM_MODE");
  const char *how = getenv("JB200_ATTACH");
  int rc;
  if (jb200_api_load(&g_api, (void *)&jb200_attach) != 0) return -1;
  if (jb200_dnn_from_blob(b, &g_dd) == 0) {
    rc = g_api.dnn_create(&g_dd, 0, &g_dnn);
  } else {
    if (jb200_gmm_from_blob(b, &g_gd) != 0) { jlog("ERROR: jb200: no acoustic model in the flattened blob\n"); return -1; }
    g_state_off = g_gd.state_off;
    rc = g_api.gmm_create(&g_gd, 0, (mode && strcmp(mode, "fast") == 0) ? JB200_GMM_FAST : JB200_GMM_EXACT, &g_gmm);
  }
  if (rc != 0) { jlog("ERROR: jb200: cannot create the GPU scorer: %s\n", g_api.last_error()); return -1; }
  if (how && strcmp(how, "calcmix") == 0) {
    /* only the -gprune jb200 surface: the host keeps calling outprob_state -> calc_mix -> calcmix() */
    callback_add(recog, CALLBACK_EVENT_PASS1_BEGIN, on_pass1_begin_calcmix, NULL);
    jlog("STAT: jb200: GPU Gaussian scoring attached behind the calcmix hook (-gprune jb200)\n");
    return 0;
  }
  callback_add(recog, CALLBACK_EVENT_PASS1_BEGIN, on_pass1_begin, NULL);
  jlog("STAT: jb200: GPU acoustic scoring attached (%s)\n", g_dnn ? "DNN, tensor cores" : "GMM");
  return 0;
}

/* ---------------------------------------------------------------- calcmix hook set (-gprune jb200) */
void calcmix_get_optname(char *buf, int buflen) { strncpy(buf, "jb200", buflen); }

boolean calcmix_init(HMMWork *wrk) {
  /* same work-area contract as gprune_none_init (libsent/src/phmm/gprune_none.c:92-103) */
  wrk->OP_calced_maxnum = wrk->OP_hmminfo->maxmixturenum * wrk->OP_nstream;
  wrk->OP_calced_score = (LOGPROB *)malloc(sizeof(LOGPROB) * wrk->OP_calced_maxnum);
  wrk->OP_calced_id = (int *)malloc(sizeof(int) * wrk->OP_calced_maxnum);
  wrk->OP_gprune_num = wrk->OP_calced_maxnum;
  return TRUE;
}

void calcmix_free(HMMWork *wrk) { free(wrk->OP_calced_score); free(wrk->OP_calced_id); }

void calcmix(HMMWork *wrk, HTK_HMM_Dens **g, int num, int *last_id, int lnum) {
  int i, base;
  if (g_gmm == NULL) { j_internal_error("jb200 calcmix: the GPU scorer is not attached (set JB200_ATTACH=1)\n"); return; }
  if (g_gauss == NULL) g_gauss = (float *)malloc(sizeof(float) * (g_gd.n_gauss + 1));
  if (wrk->OP_time != g_gauss_time) {
    if (g_api.gmm_gauss_host(g_gmm, wrk->OP_vec, g_gauss) != 0) j_internal_error("jb200 calcmix: %s\n", g_api.last_error());
    g_gauss_time = wrk->OP_time;
  }
  base = g_state_off[wrk->OP_state_id];
  for (i = 0; i < num; i++) { wrk->OP_calced_score[i] = g_gauss[base + i]; wrk->OP_calced_id[i] = i; }
  wrk->OP_calced_num = num;
}
