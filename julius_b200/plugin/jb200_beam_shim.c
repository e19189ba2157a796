/* jb200_beam_shim.c -- link-time replacement of libjulius/src/beam.c's entry points.
 *
 * pass1.c calls the pass-1 beam by name (libjulius/src/pass1.c:234,242,503,409/567;
 * libjulius/include/julius/extern.h:56-61).  Linking libjulius with this object INSTEAD of beam.o
 * turns the stock Julius host into a front for the GPU path, jconf surface unchanged:
 *
 *   get_back_trellis_init(param, r)      beam.c:1825  bt_prepare and the per-utterance host state, as the original
 *   get_back_trellis_proceed(t, ...)     beam.c:2663  nothing to do per frame; returns TRUE, no interim result
 *   get_back_trellis_end(param, r)       beam.c:3052  param now holds ALL frames of the input in both of the host's
 *                                                    modes -- buffered (pass1.c:220-254 calls _init with the whole
 *                                                    utterance) and real-time (realtime-1stpass.c:681 calls _init with
 *                                                    ONE frame and grows param as audio arrives) -- so this is where
 *                                                    the utterance is scored and decoded on the GPU
 *                                                    (jb200_decode_batch_host) and r->backtrellis is materialised from
 *                                                    the GPU's atoms through bt_new/bt_store (backtrellis.c:154,190)
 *   finalize_1st_pass(r, len)            beam.c:3133  bt_relocate_rw + bt_sort_rw, then publishes the
 *                                                    pass-1 best exactly where find_1pass_result does
 *                                                    (beam.c:497-512)
 *   fsbeam_free(d)                       beam.c:3180
 * Restrictions (checked, fail loudly): N-gram LM, no short-pause segmentation, feature-vector input at least as wide
 * as the model's.  Progressive (interim) output is not produced: pass 1 runs when the input is complete.
 * (normal and multipath trees both run on the device).  The models are flattened on first use with the same code as the plugin.
 */
#include <julius/juliuslib.h>
#include "jb200_model.h"
#include "jb200_dl.h"

extern int jb200_flatten(PROCESS_AM *am, RecogProcess *r, jb200_blob *b);   /* jb200_export.c */

typedef struct {
  RecogProcess *r;
  jb200_blob blob;
  jb200_gmm_desc gd; jb200_dnn_desc dd; jb200_tree_desc td;
  jb200_gmm *gmm; jb200_dnn *dnn; jb200_decoder *dec;
  int max_frames;
  boolean ok;            /* last decode succeeded */
} Shim;

static jb200_api g_api;
static int g_api_loaded = 0;
static Shim g_shim[8];
static int g_nshim = 0;

static Shim *shim_for(RecogProcess *r, int frames) {
  int i, rc;
  Shim *s = NULL;
  const char *mode = getenv("JB200_GMM_MODE");
  for (i = 0; i < g_nshim; i++) if (g_shim[i].r == r) s = &g_shim[i];
  if (s && frames <= s->max_frames) return s;
  if (!g_api_loaded) { if (jb200_api_load(&g_api, (void *)&shim_for) != 0) return NULL; g_api_loaded = 1; }
  if (s == NULL) {
    if (g_nshim >= 8) { jlog("ERROR: jb200: too many recognition instances\n"); return NULL; }
    if (r->lmtype != LM_PROB || r->config->successive.enabled) {
      jlog("ERROR: jb200: the GPU beam supports N-gram, non-segmented decoding only\n");
      return NULL;
    }
    s = &g_shim[g_nshim];
    memset(s, 0, sizeof(*s));
    s->r = r;
    jb200_blob_init(&s->blob);
    if (jb200_flatten(r->am, r, &s->blob) != 0) return NULL;
    if (jb200_tree_from_blob(&s->blob, &s->td) != 0) { jlog("ERROR: jb200: no lexicon tree in the flattened model\n"); return NULL; }
    if (jb200_dnn_from_blob(&s->blob, &s->dd) == 0) {
      /* DNN-HMM: a Gaussian-free scorer carries the state / cd-set layout */
      memset(&s->gd, 0, sizeof(s->gd));
      s->gd.n_states = jb200_blob_get_i(&s->blob, "gmm.n_states", 0);
      s->gd.iwcd_method = jb200_blob_get_i(&s->blob, "am.iwcd_method", JB200_IWCD_NBEST);
      s->gd.iwcd_nbest = jb200_blob_get_i(&s->blob, "am.iwcd_nbest", 3);
      s->gd.n_cdsets = jb200_blob_get_i(&s->blob, "am.n_cdsets", 0);
      s->gd.n_cdset_states = jb200_blob_get_i(&s->blob, "am.n_cdset_states", 0);
      s->gd.cd_off = (const int32_t *)jb200_blob_ptr(&s->blob, "am.cd_off", NULL);
      s->gd.cd_states = (const int32_t *)jb200_blob_ptr(&s->blob, "am.cd_states", NULL);
      s->gd.dim = s->dd.in_dim;
      rc = g_api.dnn_create(&s->dd, 0, &s->dnn);
      if (rc != 0) { jlog("ERROR: jb200: %s\n", g_api.last_error()); return NULL; }
    } else if (jb200_gmm_from_blob(&s->blob, &s->gd) != 0) { jlog("ERROR: jb200: no acoustic model\n"); return NULL; }
    rc = g_api.gmm_create(&s->gd, 0, (mode && strcmp(mode, "fast") == 0) ? JB200_GMM_FAST : JB200_GMM_EXACT, &s->gmm);
    if (rc != 0) { jlog("ERROR: jb200: %s\n", g_api.last_error()); return NULL; }
    g_nshim++;
    jlog("STAT: jb200: GPU pass-1 beam attached to %02d %s\n", r->config->id, r->config->name);
  }
  /* (re)create the decoder for the longest utterance seen so far: the old one (device work areas, pinned host
   * buffers) is released first, and the capacity is recorded only once the new one exists */
  {
    const int want = frames < 4096 ? 4096 : frames + frames / 2;
    if (s->dec) { g_api.decoder_destroy(s->dec); s->dec = NULL; s->max_frames = 0; }
    rc = g_api.decoder_create(&s->td, s->gmm, 1, want, &s->dec);
    if (rc == 0 && s->dnn) rc = g_api.decoder_attach_dnn(s->dec, s->dnn);
    if (rc != 0) {
      jlog("ERROR: jb200: %s\n", g_api.last_error());
      if (s->dec) { g_api.decoder_destroy(s->dec); s->dec = NULL; }
      return NULL;
    }
    s->max_frames = want;
  }
  return s;
}

boolean get_back_trellis_init(HTK_Param *param, RecogProcess *r) {
  Shim *s;
  bt_prepare(r->backtrellis);
  r->pass1.bos.wid = WORD_INVALID;
  r->pass1.bos.begintime = r->pass1.bos.endtime = -1;
  /* host-side state that init_nodescore resets per utterance and pass 2 relies on (beam.c:1595):
   * the per-node triphone caches of outprob_style (bt_discount_pescore and the stack decoder read them) */
  outprob_style_cache_init(r->wchmm);
  r->config->output.progout_interval_frame = (int)((float)r->config->output.progout_interval / ((float)param->header.wshift / 10000.0));
  s = shim_for(r, 1);
  if (s == NULL) return FALSE;
  s->ok = FALSE;
  if (param->is_outprob) { jlog("ERROR: jb200: outprob-vector input is not supported by the GPU beam shim\n"); return FALSE; }
  if (param->veclen < s->gd.dim) {
    jlog("ERROR: jb200: input vectors have %d components, the acoustic model takes %d\n", (int)param->veclen, s->gd.dim);
    return FALSE;
  }
  return TRUE;
}

boolean get_back_trellis_proceed(int t, HTK_Param *param, RecogProcess *r, boolean final_for_multipath) {
  r->have_interim = FALSE;
  return TRUE;
}

void get_back_trellis_end(HTK_Param *param, RecogProcess *r) {
  const jb200_utt_result *u; const jb200_atom *a; const int32_t *w;
  TRELLIS_ATOM **idx;
  const int T = param->samplenum;
  Shim *s = shim_for(r, T);
  int i, t, D;
  int32_t off[2];
  float *in;
  if (s == NULL) return;
  s->ok = FALSE;
  if (T < 1 || param->is_outprob || param->veclen < s->gd.dim) return;       /* refused at _init already */
  D = s->gd.dim;
  in = (float *)malloc(sizeof(float) * (size_t)T * D);
  if (in == NULL) { jlog("ERROR: jb200: out of memory\n"); return; }
  for (t = 0; t < T; t++) memcpy(in + (size_t)t * D, param->parvec[t], sizeof(float) * D);
  off[0] = 0; off[1] = T;
  if (g_api.decode_batch_host(s->dec, in, off, 1) != 0) { jlog("ERROR: jb200: %s\n", g_api.last_error()); free(in); return; }
  free(in);
  s->ok = TRUE;
  if (g_api.decoder_results(s->dec, &u, &a, &w) != 0) return;
  if (u->overflow) { jlog("ERROR: jb200: device work area overflow (code %d); pass 1 result dropped\n", u->overflow); return; }
  a += u->atom_offset;
  idx = (TRELLIS_ATOM **)malloc(sizeof(void *) * (u->n_atoms + 1));
  for (i = 0; i < u->n_atoms; i++) {
    TRELLIS_ATOM *tre = bt_new(r->backtrellis);
    tre->wid = (WORD_ID)a[i].wid;
    tre->begintime = (short)a[i].begintime; tre->endtime = (short)a[i].endtime;
    tre->backscore = a[i].backscore; tre->lscore = a[i].lscore;
    tre->dfa_state = -1;
    tre->last_tre = (a[i].last < 0) ? &(r->pass1.bos) : idx[a[i].last];
    bt_store(r->backtrellis, tre);
    idx[i] = tre;
  }
  free(idx);
}

void finalize_1st_pass(RecogProcess *r, int len) {
  const jb200_utt_result *u; const jb200_atom *a; const int32_t *w;
  BACKTRELLIS *bt = r->backtrellis;
  Shim *s = shim_for(r, 0);
  int i;
  bt->framelen = len;
  bt_relocate_rw(bt);
  bt_sort_rw(bt);
  if (bt->num == NULL || s == NULL || !s->ok || g_api.decoder_results(s->dec, &u, &a, &w) != 0) {
    if (bt->framelen > 0) jlog("WARNING: %02d %s: input processed, but no survived word found\n", r->config->id, r->config->name);
    r->result.status = J_RESULT_STATUS_FAIL;
    return;
  }
  if (u->status != 0) {
    jlog("WARNING: %02d %s: no tail silence word survived on the last frame, search failed\n", r->config->id, r->config->name);
    r->result.status = J_RESULT_STATUS_FAIL;
    return;
  }
  w += u->word_offset;
  /* what find_1pass_result publishes (beam.c:497-517) */
  r->result.status = J_RESULT_STATUS_SUCCESS;
  r->result.num_frame = len;
  for (i = 0; i < u->n_words; i++) r->result.pass1.word[i] = (WORD_ID)w[i];
  r->result.pass1.word_num = u->n_words;
  r->result.pass1.score = u->score;
  {
    /* total LM score along the best path = sum of lscore of its atoms (trace_backptr, beam.c:253-301) */
    LOGPROB lsum = 0.0; int k, last_time = len - 1, best = -1;
    a += u->atom_offset;
    for (k = u->n_atoms - 1; k >= 0 && best < 0; k--)
      if (a[k].wid == (int)r->lm->winfo->tail_silwid && a[k].backscore == u->score) best = k;
    (void)last_time;
    for (k = best; k >= 0; k = a[k].last) { lsum += a[k].lscore; if (a[k].begintime <= 0) break; }
    r->result.pass1.score_lm = lsum;
    r->result.pass1.score_am = u->score - lsum;
  }
  for (i = 0; i < u->n_words; i++) r->pass1_wseq[i] = (WORD_ID)w[i];
  r->pass1_wnum = u->n_words;
  r->pass1_score = u->score;
}

void fsbeam_free(FSBeam *d) {
  int i;
  if (d->pausemodelnames != NULL) { free(d->pausemodelnames); free(d->pausemodel); }
  if (d->boslist != NULL) free(d->boslist);
  /* release the device side of the recognition instance this work area belongs to */
  for (i = 0; i < g_nshim; i++) {
    Shim *s = &g_shim[i];
    if (s->r == NULL || &(s->r->pass1) != d) continue;
    if (s->dec) { g_api.decoder_destroy(s->dec); s->dec = NULL; }
    if (s->dnn) { g_api.dnn_destroy(s->dnn); s->dnn = NULL; }
    if (s->gmm) { g_api.gmm_destroy(s->gmm); s->gmm = NULL; }
    jb200_blob_free(&s->blob);
    s->r = NULL; s->max_frames = 0; s->ok = FALSE;
  }
}
