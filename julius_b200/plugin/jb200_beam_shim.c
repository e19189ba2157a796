/* jb200_beam_shim.c -- link-time replacement of libjulius/src/beam.c's entry points.
 *
 * pass1.c calls the pass-1 beam by name (libjulius/src/pass1.c:234,242,503,409/567;
 * libjulius/include/julius/extern.h:56-61).  Linking libjulius with this object INSTEAD of beam.o
 * turns the stock Julius host into a front for the GPU path, jconf surface unchanged:
 *
 *   get_back_trellis_init(param, r)      beam.c:1825  bt_prepare and the per-utterance host state, as the original
 *   get_back_trellis_proceed(t, ...)     beam.c:2663  two modes.  Buffered (default for file input): nothing per frame.
 *                                                    Frame-synchronous (real-time input, -progout, or JB200_STREAM=1):
 *                                                    frame t has just arrived in param (realtime-1stpass.c:681 calls
 *                                                    _init with ONE frame and grows param as audio comes in); the
 *                                                    frames not yet decoded are fed to the device stream every
 *                                                    JB200_STREAM_FRAMES frames (default 10 = 100 ms), and always when
 *                                                    the host is due an interim result (beam.c:2983-2993: every
 *                                                    progout_interval_frame frames r->result.pass1 gets the best word
 *                                                    sequence so far and have_interim is raised); returns FALSE once
 *                                                    the beam ran empty (beam.c:3012-3015)
 *   get_back_trellis_end(param, r)       beam.c:3052  param now holds ALL frames of the input.  Buffered mode: this is
 *                                                    where the utterance is scored and decoded on the GPU
 *                                                    (jb200_decode_batch_host); frame-synchronous mode: the remaining
 *                                                    frames go to the stream together with the end-of-utterance mark.
 *                                                    Either way r->backtrellis is materialised from the GPU's atoms
 *                                                    through bt_new/bt_store (backtrellis.c:154,190)
 *   finalize_1st_pass(r, len)            beam.c:3133  bt_relocate_rw + bt_sort_rw, then publishes the
 *                                                    pass-1 best exactly where find_1pass_result does
 *                                                    (beam.c:497-512)
 *   fsbeam_free(d)                       beam.c:3180
 * Restrictions (checked, fail loudly): N-gram LM, no short-pause segmentation, feature-vector input at least as wide
 * as the model's.
 * (normal and multipath trees both run on the device).  The models are flattened on first use with the same code as the plugin.
 */
#include <julius/juliuslib.h>
#include "jb200_model.h"
#include "jb200_dl.h"
#include <time.h>

static double shim_now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

extern int jb200_flatten(PROCESS_AM *am, RecogProcess *r, jb200_blob *b);   /* jb200_export.c */

/* one decoded utterance, copied out of the decoder (decode-ahead cache, and the current utterance) */
typedef struct {
  unsigned long long hash;   /* of the feature vectors it was decoded from */
  int n_frames;
  jb200_utt_result u;
  jb200_atom *atoms;         /* [u.n_atoms] */
  int32_t *words;            /* [u.n_words] */
} ShimResult;

typedef struct {
  RecogProcess *r;
  jb200_blob blob;
  jb200_gmm_desc gd; jb200_dnn_desc dd; jb200_tree_desc td;
  jb200_gmm *gmm; jb200_dnn *dnn; jb200_decoder *dec;
  int max_frames, max_utts;
  boolean ok;            /* last decode succeeded */
  /* frame-synchronous mode */
  boolean streaming;     /* this utterance runs on a device stream */
  int fed;               /* frames handed to the stream so far */
  int stream_frames;     /* feed granularity */
  float *stage; int stage_cap;   /* packing buffer for one feed */
  ShimResult cur;        /* result of the utterance being finished */
  /* decode-ahead over a file list (JB200_FILELIST = the list given to -filelist, JB200_AHEAD = how many files a batch) */
  char **files; int n_files, next_file;   /* next_file: index of the utterance the host will finish next */
  ShimResult *ahead; int n_ahead, ahead_first;
  long n_from_cache, n_single;
} Shim;

static jb200_api g_api;
static int g_api_loaded = 0;
static Shim g_shim[8];
static int g_nshim = 0;

static Shim *shim_for2(RecogProcess *r, int frames, int utts);
static Shim *shim_for(RecogProcess *r, int frames) { return shim_for2(r, frames, 1); }
static Shim *shim_for2(RecogProcess *r, int frames, int utts) {
  int i, rc;
  Shim *s = NULL;
  const char *mode = getenv("JB200_GMM_MODE");
  for (i = 0; i < g_nshim; i++) if (g_shim[i].r == r) s = &g_shim[i];
  if (s && frames <= s->max_frames && utts <= s->max_utts) return s;
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
    int want = frames < 4096 ? 4096 : frames + frames / 2;
    const int want_utts = utts > s->max_utts ? utts : (s->max_utts > 0 ? s->max_utts : 1);
    if (want < s->max_frames) want = s->max_frames;
    if (s->dec) { g_api.decoder_destroy(s->dec); s->dec = NULL; s->max_frames = 0; s->max_utts = 0; }
    rc = g_api.decoder_create(&s->td, s->gmm, want_utts, want, &s->dec);
    if (rc == 0 && s->dnn) rc = g_api.decoder_attach_dnn(s->dec, s->dnn);
    if (rc != 0) {
      jlog("ERROR: jb200: %s\n", g_api.last_error());
      if (s->dec) { g_api.decoder_destroy(s->dec); s->dec = NULL; }
      return NULL;
    }
    s->max_frames = want; s->max_utts = want_utts;
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
  /* frame-synchronous decoding when the input is live, when the host wants interim results, or on request */
  {
    const char *e = getenv("JB200_STREAM"), *f = getenv("JB200_STREAM_FRAMES");
    /* live input: realtime-1stpass.c:681-682 hands _init the first frame alone (a RecogProcess has no way to ask its
     * Recog for decodeopt.realtime_flag); a buffered one-frame input takes the same route, which is equally right */
    const boolean live = (param->samplenum <= 1);
    const int shift = (r->am != NULL && r->am->config != NULL) ? r->am->config->analysis.para.frameshift : 0;
    s->streaming = (e != NULL) ? (atoi(e) != 0) : (live || r->config->output.progout_flag);
    s->stream_frames = (f != NULL && atoi(f) > 0) ? atoi(f) : 10;
    s->fed = 0;
    if (s->streaming) {
      /* live input has no length yet: room for the longest input the host accepts (MAXSPEECHLEN samples) */
      int cap = (shift > 0) ? MAXSPEECHLEN / shift + 16 : 4096;
      if (cap < param->samplenum) cap = param->samplenum;
      s = shim_for(r, cap);
      if (s == NULL) return FALSE;
      if (g_api.stream_open(s->dec, 1) != 0) { jlog("ERROR: jb200: %s\n", g_api.last_error()); return FALSE; }
    }
  }
  return TRUE;
}

/* hand frames [s->fed, upto) of param to the device stream */
static boolean stream_push(Shim *s, HTK_Param *param, int upto, boolean last, boolean interim) {
  const int D = s->gd.dim, n = upto - s->fed;
  int32_t n_new = n; uint8_t fin = last ? 1 : 0;
  int t;
  if (n < 0) return FALSE;
  if (n > s->stage_cap) {
    free(s->stage);
    s->stage_cap = n < 64 ? 64 : n;
    s->stage = (float *)malloc(sizeof(float) * (size_t)s->stage_cap * D);
    if (s->stage == NULL) { s->stage_cap = 0; jlog("ERROR: jb200: out of memory\n"); return FALSE; }
  }
  for (t = 0; t < n; t++) memcpy(s->stage + (size_t)t * D, param->parvec[s->fed + t], sizeof(float) * D);
  if (g_api.stream_feed_host(s->dec, s->stage, &n_new, &fin, interim ? 1 : 0) != 0) { jlog("ERROR: jb200: %s\n", g_api.last_error()); return FALSE; }
  s->fed = upto;
  return TRUE;
}

boolean get_back_trellis_proceed(int t, HTK_Param *param, RecogProcess *r, boolean final_for_multipath) {
  Shim *s = shim_for(r, 0);
  boolean want_interim;
  r->have_interim = FALSE;
  if (s == NULL || !s->streaming || final_for_multipath) return TRUE;
  /* beam.c:2983-2993: after frame t the best path ending at t-1 is published every progout_interval_frame frames */
  want_interim = (t > 0 && r->config->output.progout_flag && r->config->output.progout_interval_frame > 0 &&
                  ((t - 1) % r->config->output.progout_interval_frame) == 0);
  if (t + 1 - s->fed < s->stream_frames && !want_interim) return TRUE;
  if (!stream_push(s, param, t + 1, FALSE, want_interim)) return FALSE;
  {
    int32_t done = 0, alive = 1, ended = 0;
    g_api.stream_status(s->dec, 0, &done, &alive, &ended);
    if (!alive) {
      jlog("ERROR: get_back_trellis_proceed: %02d %s: frame %d: no nodes left in beam, now terminates search\n", r->config->id, r->config->name, t);
      return FALSE;
    }
  }
  if (want_interim) {
    int32_t words[MAXSEQNUM], nw = 0, frame = -1; float score = LOG_ZERO; int i;
    if (g_api.stream_partial(s->dec, 0, words, MAXSEQNUM, &nw, &score, &frame) == 0) {
      /* what bt_current_max leaves in r->result (beam.c:898-920) */
      r->have_interim = TRUE;
      r->result.status = J_RESULT_STATUS_SUCCESS;
      r->result.num_frame = t - 1;
      r->result.pass1.word_num = nw;
      for (i = 0; i < nw; i++) r->result.pass1.word[i] = (WORD_ID)words[i];
      if (nw > 0) { r->result.pass1.score = score; r->result.pass1.score_am = score; r->result.pass1.score_lm = 0.0; }
      if (getenv("JB200_SHIM_VERBOSE")) { printf("JB200_SHIM interim t=%d words=%d score=%f\n", t, (int)nw, score); fflush(stdout); }
    }
  }
  return TRUE;
}

/* ---- results: copies that outlive the decoder's buffers -------------------------------------------------------- */
static void result_free(ShimResult *x) { free(x->atoms); free(x->words); memset(x, 0, sizeof(*x)); }

static int result_copy(ShimResult *x, const jb200_utt_result *u, const jb200_atom *atoms, const int32_t *words) {
  result_free(x);
  x->u = *u;
  x->atoms = (jb200_atom *)malloc(sizeof(jb200_atom) * (size_t)(u->n_atoms > 0 ? u->n_atoms : 1));
  x->words = (int32_t *)malloc(sizeof(int32_t) * (size_t)(u->n_words > 0 ? u->n_words : 1));
  if (!x->atoms || !x->words) { result_free(x); return -1; }
  if (u->n_atoms > 0) memcpy(x->atoms, atoms + u->atom_offset, sizeof(jb200_atom) * (size_t)u->n_atoms);
  if (u->n_words > 0) memcpy(x->words, words + u->word_offset, sizeof(int32_t) * (size_t)u->n_words);
  x->u.atom_offset = 0; x->u.word_offset = 0;
  return 0;
}

static unsigned long long feat_hash(const float *x, size_t n) {      /* FNV-1a over the bit patterns */
  const unsigned char *b = (const unsigned char *)x;
  unsigned long long h = 1469598103934665603ULL;
  size_t i;
  for (i = 0; i < n * sizeof(float); i++) { h ^= b[i]; h *= 1099511628211ULL; }
  return h;
}

/* ---- decode-ahead: the host hands over one utterance at a time (pass1.c:220-254), one utterance occupies one of several
 * hundred resident thread blocks.  With JB200_FILELIST = the list the host reads its HTK parameter files from, the shim
 * reads the next JB200_AHEAD files itself, decodes them in ONE batch, and answers the host's following utterances from
 * the cache -- but only when the vectors the host presents hash to what was decoded (any host-side processing of the
 * input, or a list that does not match, silently falls back to the one-utterance path). */
static float *read_htk_param(const char *fn, int want_dim, int *n_frames) {
  FILE *fp = fopen(fn, "rb");
  unsigned char h[12];
  unsigned int ns, ssize;
  float *x; size_t i, n;
  if (!fp) return NULL;
  if (fread(h, 1, 12, fp) != 12) { fclose(fp); return NULL; }
  ns = ((unsigned)h[0] << 24) | ((unsigned)h[1] << 16) | ((unsigned)h[2] << 8) | h[3];
  ssize = ((unsigned)h[8] << 8) | h[9];
  if (ns < 1 || ns > 32767 || ssize != (unsigned)want_dim * 4u) { fclose(fp); return NULL; }
  n = (size_t)ns * want_dim;
  x = (float *)malloc(sizeof(float) * n);
  if (!x || fread(x, 4, n, fp) != n) { free(x); fclose(fp); return NULL; }
  fclose(fp);
  for (i = 0; i < n; i++) {                                           /* big-endian floats (rdparam.c:83-187) */
    unsigned char *b = (unsigned char *)(x + i), t;
    t = b[0]; b[0] = b[3]; b[3] = t; t = b[1]; b[1] = b[2]; b[2] = t;
  }
  *n_frames = (int)ns;
  return x;
}

static void load_filelist(Shim *s) {
  const char *fn = getenv("JB200_FILELIST");
  char line[4096]; FILE *fp;
  s->n_files = 0; s->files = NULL;
  if (!fn || !(fp = fopen(fn, "r"))) return;
  while (fgets(line, sizeof(line), fp)) {
    size_t L = strlen(line);
    while (L > 0 && (line[L - 1] == '\n' || line[L - 1] == '\r' || line[L - 1] == ' ')) line[--L] = '\0';
    if (L == 0 || line[0] == '#') continue;
    s->files = (char **)realloc(s->files, sizeof(char *) * (size_t)(s->n_files + 1));
    s->files[s->n_files++] = strdup(line);
  }
  fclose(fp);
  jlog("STAT: jb200: decode-ahead over %d files of %s\n", s->n_files, fn);
}

static void ahead_clear(Shim *s) {
  int i;
  for (i = 0; i < s->n_ahead; i++) result_free(&s->ahead[i]);
  free(s->ahead); s->ahead = NULL; s->n_ahead = 0;
}

/* decode files [first, first+k) in one batch; leaves whatever could be decoded in the cache */
static void ahead_fill(Shim *s, int first) {
  const char *e = getenv("JB200_AHEAD");
  int k = e ? atoi(e) : 32, i, n = 0, total = 0, D = s->gd.dim;
  float **xs; int *T; float *cat; int32_t *off;
  const jb200_utt_result *u; const jb200_atom *a; const int32_t *w;
  Shim *s2;
  const double t0 = shim_now(); double t1 = t0, t2 = t0, t3 = t0;
  ahead_clear(s);
  if (k < 2) return;
  if (first + k > s->n_files) k = s->n_files - first;
  if (k < 2) return;
  xs = (float **)calloc((size_t)k, sizeof(float *)); T = (int *)calloc((size_t)k, sizeof(int));
  for (i = 0; i < k; i++) {
    xs[i] = read_htk_param(s->files[first + i], D, &T[i]);
    if (!xs[i]) break;
    total += T[i]; n++;
  }
  t1 = shim_now();
  if (n >= 2 && (s2 = shim_for2(s->r, total, n)) != NULL) {
    t2 = shim_now();
    cat = (float *)malloc(sizeof(float) * (size_t)total * D);
    off = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
    off[0] = 0;
    for (i = 0; i < n; i++) { memcpy(cat + (size_t)off[i] * D, xs[i], sizeof(float) * (size_t)T[i] * D); off[i + 1] = off[i] + T[i]; }
    if (g_api.decode_batch_host(s->dec, cat, off, n) == 0 && g_api.decoder_results(s->dec, &u, &a, &w) == 0) {
      t3 = shim_now();
      s->ahead = (ShimResult *)calloc((size_t)n, sizeof(ShimResult));
      s->n_ahead = n; s->ahead_first = first;
      if (getenv("JB200_SHIM_VERBOSE")) { printf("JB200_SHIM batch first=%d n=%d frames=%d read=%.3fs decoder=%.3fs decode=%.3fs\n", first, n, total, t1 - t0, t2 - t1, t3 - t2); fflush(stdout); }
      for (i = 0; i < n; i++) {
        result_copy(&s->ahead[i], &u[i], a, w);
        s->ahead[i].hash = feat_hash(xs[i], (size_t)T[i] * D);
        s->ahead[i].n_frames = T[i];
      }
    } else jlog("WARNING: jb200: decode-ahead batch failed (%s); continuing one utterance at a time\n", g_api.last_error());
    free(cat); free(off);
  }
  for (i = 0; i < k; i++) free(xs[i]);
  free(xs); free(T);
}

void get_back_trellis_end(HTK_Param *param, RecogProcess *r) {
  TRELLIS_ATOM **idx;
  const int T = param->samplenum;
  Shim *s = shim_for(r, T);
  int i, t, D, utt;
  float *in;
  unsigned long long h;
  if (s == NULL) return;
  s->ok = FALSE;
  if (T < 1 || param->is_outprob || param->veclen < s->gd.dim) return;       /* refused at _init already */
  if (s->streaming) {
    /* frame-synchronous mode: the rest of the input and the end-of-utterance mark */
    const jb200_utt_result *u; const jb200_atom *a; const int32_t *w;
    if (!stream_push(s, param, T, TRUE, FALSE)) return;
    if (g_api.stream_result(s->dec, 0, &u, &a, &w) != 0 || result_copy(&s->cur, u, a, w) != 0) { jlog("ERROR: jb200: %s\n", g_api.last_error()); return; }
    s->ok = TRUE; s->next_file++;
    goto materialise;
  }
  D = s->gd.dim;
  in = (float *)malloc(sizeof(float) * (size_t)T * D);
  if (in == NULL) { jlog("ERROR: jb200: out of memory\n"); return; }
  for (t = 0; t < T; t++) memcpy(in + (size_t)t * D, param->parvec[t], sizeof(float) * D);
  /* answered by the decode-ahead cache? */
  utt = s->next_file++;
  if (s->files == NULL && s->n_files == 0 && getenv("JB200_FILELIST")) { load_filelist(s); if (s->n_files == 0) s->n_files = -1; }
  if (s->n_files > 0 && utt < s->n_files) {
    if (!(s->n_ahead > 0 && utt >= s->ahead_first && utt < s->ahead_first + s->n_ahead)) ahead_fill(s, utt);
    if (s->n_ahead > 0 && utt >= s->ahead_first && utt < s->ahead_first + s->n_ahead) {
      ShimResult *c = &s->ahead[utt - s->ahead_first];
      h = feat_hash(in, (size_t)T * D);
      if (c->atoms != NULL && c->n_frames == T && c->hash == h && result_copy(&s->cur, &c->u, c->atoms, c->words) == 0) {
        s->ok = TRUE; s->n_from_cache++;
        if (getenv("JB200_SHIM_VERBOSE")) { printf("JB200_SHIM utt=%d from_cache\n", utt); fflush(stdout); }
      }
    }
  }
  if (!s->ok) {
    const jb200_utt_result *u; const jb200_atom *a; const int32_t *w;
    int32_t off[2];
    off[0] = 0; off[1] = T;
    if (g_api.decode_batch_host(s->dec, in, off, 1) != 0 || g_api.decoder_results(s->dec, &u, &a, &w) != 0 ||
        result_copy(&s->cur, u, a, w) != 0) { jlog("ERROR: jb200: %s\n", g_api.last_error()); free(in); return; }
    s->ok = TRUE; s->n_single++;
  }
  free(in);
materialise:
  if (s->cur.u.overflow) { jlog("ERROR: jb200: device work area overflow (code %d); pass 1 result dropped\n", s->cur.u.overflow); s->ok = FALSE; return; }
  idx = (TRELLIS_ATOM **)malloc(sizeof(void *) * (size_t)(s->cur.u.n_atoms + 1));
  for (i = 0; i < s->cur.u.n_atoms; i++) {
    const jb200_atom *a = s->cur.atoms;
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
  if (bt->num == NULL || s == NULL || !s->ok) {
    if (bt->framelen > 0) jlog("WARNING: %02d %s: input processed, but no survived word found\n", r->config->id, r->config->name);
    r->result.status = J_RESULT_STATUS_FAIL;
    return;
  }
  u = &s->cur.u; a = s->cur.atoms; w = s->cur.words;
  if (u->status != 0) {
    jlog("WARNING: %02d %s: no tail silence word survived on the last frame, search failed\n", r->config->id, r->config->name);
    r->result.status = J_RESULT_STATUS_FAIL;
    return;
  }
  /* what find_1pass_result publishes (beam.c:497-517) */
  r->result.status = J_RESULT_STATUS_SUCCESS;
  r->result.num_frame = len;
  for (i = 0; i < u->n_words; i++) r->result.pass1.word[i] = (WORD_ID)w[i];
  r->result.pass1.word_num = u->n_words;
  r->result.pass1.score = u->score;
  {
    /* total LM score along the best path = sum of lscore of its atoms (trace_backptr, beam.c:253-301) */
    LOGPROB lsum = 0.0; int k, best = -1;
    for (k = u->n_atoms - 1; k >= 0 && best < 0; k--)
      if (a[k].wid == (int)r->lm->winfo->tail_silwid && a[k].backscore == u->score) best = k;
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
    if (s->n_files > 0) jlog("STAT: jb200: %ld utterances answered from decode-ahead batches, %ld decoded singly\n", s->n_from_cache, s->n_single);
    ahead_clear(s); result_free(&s->cur);
    free(s->stage); s->stage = NULL; s->stage_cap = 0;
    jb200_blob_free(&s->blob);
    s->r = NULL; s->max_frames = 0; s->ok = FALSE;
  }
}
