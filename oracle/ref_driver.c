/* oracle/ref_driver.c -- TEST INFRASTRUCTURE.  Not part of the product.
 *
 * A small driver over the UNMODIFIED reference's JuliusLib API
 * (libjulius/include/julius/juliuslib.h; pattern of julius-simple/julius-simple.c:226-362)
 * that runs recognition exactly as `julius` would for the given jconf-style
 * arguments and dumps, per input utterance, what the parity harness needs:
 *
 *   - the [T x S] state log10-likelihood matrix out of wrk->outprob_cache
 *     (the same numbers `-outprobout` writes, libsent/src/phmm/outprob.c:440-485)
 *   - the finalized word trellis r->backtrellis->rw[t][i]  (trellis.h:28-53,
 *     after bt_relocate_rw/bt_sort_rw, beam.c:3133-3162)
 *   - the pass-1 best word sequence and score (beam.c:497-512)
 *   - optionally (JREF_TOKENS=1) the surviving token set of every frame
 *     (beam.h:35-45, FSBeam tlist/tindex/n_start/n_end) for frame-level debugging
 *   - decode wall time between CALLBACK_EVENT_PASS1_BEGIN and _END (CPU baseline)
 *
 * usage: jref -dump out.jrf [julius options ...]
 * The dump is little-endian, see julius_b200/refdump.py for the reader.
 */
#include <julius/juliuslib.h>
#include <time.h>

static FILE *g_out = NULL;
static int g_tokens = 0;
static double g_t0 = 0.0;
static double g_decode_sec = 0.0;
static long g_frames = 0;
static int g_utt = 0;

static double now_sec(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void wi(int v) { fwrite(&v, 4, 1, g_out); }
static void wf(float v) { fwrite(&v, 4, 1, g_out); }

static int cmp_ptr(const void *a, const void *b) {
  const TRELLIS_ATOM *x = *(TRELLIS_ATOM * const *)a, *y = *(TRELLIS_ATOM * const *)b;
  return (x < y) ? -1 : (x > y);
}

static int atom_index(TRELLIS_ATOM **sorted, int *perm, int n, TRELLIS_ATOM *p) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    int mid = (lo + hi) / 2;
    if (sorted[mid] == p) return perm[mid];
    if (sorted[mid] < p) lo = mid + 1; else hi = mid - 1;
  }
  return -1;   /* bos (d->bos lives outside the trellis arena) */
}

static void on_pass1_begin(Recog *recog, void *dummy) { g_t0 = now_sec(); }

static void on_pass1_frame(Recog *recog, void *dummy) {
  RecogProcess *r = recog->process_list;
  FSBeam *d = &(r->pass1);
  int j, tn = d->tn;
  if (!g_tokens) return;
#ifdef JREF_NO_TOKENS
  return;
#endif
  /* record: tag, frame, count, then (node, score, last_tre wid, last_tre endtime, last_cword, last_lscore) */
  wi(0x544f4b31); /* 'TOK1' */
  wi(r->am->mfcc->f);
  wi(d->tnum[tn]);
  wi(d->n_end - d->n_start + 1);
  for (j = d->n_start; j <= d->n_end; j++) {
    TOKEN2 *tk = &(d->tlist[tn][d->tindex[tn][j]]);
    wi(tk->node); wf(tk->score);
    wi(tk->last_tre ? (int)tk->last_tre->wid : -2);
    wi(tk->last_tre ? (int)tk->last_tre->endtime : -2);
    wi((int)tk->last_cword); wf(tk->last_lscore);
  }
}

static void on_pass1_end(Recog *recog, void *dummy) {
  RecogProcess *r = recog->process_list;
  BACKTRELLIS *bt = r->backtrellis;
  HMMWork *wrk = &(r->am->hmmwrk);
  int T = r->am->mfcc->param->samplenum;
  int S = wrk->statenum;
  int t, i, s, n = 0, k;
  TRELLIS_ATOM **flat, **sorted;
  int *perm;
  double dt = now_sec() - g_t0;
  g_decode_sec += dt;
  g_frames += T;
  if (getenv("JREF_PER_UTT")) { fprintf(stdout, "JREF_UTT idx=%d frames=%d decode_sec=%.6f\n", g_utt, T, dt); fflush(stdout); }

  wi(0x4a524631); /* 'JRF1' */
  wi(g_utt++); wi(T);
  { float f = (float)dt; wf(f); }
  /* state score matrix (LOG_UNDEF where the reference never computed a state) */
  if (wrk->outprob_cache != NULL && wrk->outprob_allocframenum >= T && !r->am->mfcc->param->is_outprob) {
    wi(S);
    for (t = 0; t < T; t++) fwrite(wrk->outprob_cache[t], 4, S, g_out);
  } else {
    wi(0);
  }
  /* trellis */
  if (bt->num != NULL) for (t = 0; t < bt->framelen; t++) n += bt->num[t];
  wi(n);
  if (n > 0) {
    flat = (TRELLIS_ATOM **)malloc(sizeof(void *) * n);
    sorted = (TRELLIS_ATOM **)malloc(sizeof(void *) * n);
    perm = (int *)malloc(sizeof(int) * n);
    k = 0;
    for (t = 0; t < bt->framelen; t++) for (i = 0; i < bt->num[t]; i++) flat[k++] = bt->rw[t][i];
    /* pointer -> flat index map */
    {
      typedef struct { TRELLIS_ATOM *p; int i; } PtrIdx;
      PtrIdx *pi = (PtrIdx *)malloc(sizeof(PtrIdx) * n);
      for (k = 0; k < n; k++) { pi[k].p = flat[k]; pi[k].i = k; }
      qsort(pi, n, sizeof(PtrIdx), cmp_ptr);
      for (k = 0; k < n; k++) { sorted[k] = pi[k].p; perm[k] = pi[k].i; }
      free(pi);
    }
    for (k = 0; k < n; k++) {
      TRELLIS_ATOM *a = flat[k];
      wi((int)a->wid); wi((int)a->begintime); wi((int)a->endtime);
      wf(a->backscore); wf(a->lscore);
      wi(a->last_tre ? atom_index(sorted, perm, n, a->last_tre) : -1);
    }
    free(flat); free(sorted); free(perm);
  }
  /* pass-1 best (stored in reverse order by find_1pass_result; write as stored) */
  wi(r->result.status);
  if (r->result.status >= 0) {
    wi(r->pass1_wnum);
    for (i = 0; i < r->pass1_wnum; i++) wi((int)r->pass1_wseq[i]);
    wf(r->pass1_score);
  } else {
    wi(0); wf(0.0f);
  }
  fflush(g_out);
}

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
/* a crash inside the host libraries is reported with a symbolic backtrace (no debugger on the test boxes) */
static void on_crash(int sig) {
  void *bt[48];
  int n = backtrace(bt, 48);
  fprintf(stderr, "jref: signal %d, backtrace:\n", sig);
  backtrace_symbols_fd(bt, n, 2);
  _exit(128 + sig);
}

/* full two-pass runs (no -1pass): the final sentence hypotheses, for end-to-end host checks */
static void on_result(Recog *recog, void *dummy) {
  RecogProcess *r = recog->process_list;
  int n, i;
  fprintf(stdout, "JREF_RESULT utt=%d status=%d", g_utt - 1, r->result.status);
  if (r->result.status >= 0) {
    for (n = 0; n < r->result.sentnum; n++) {
      Sentence *st = &(r->result.sent[n]);
      union { float f; unsigned u; } sc;
      sc.f = st->score;
      fprintf(stdout, " sent%d=%08x:", n, sc.u);
      for (i = 0; i < st->word_num; i++) fprintf(stdout, "%s%d", i ? "," : "", (int)st->word[i]);
    }
  }
  fprintf(stdout, "\n");
}

/* -userlm: a deterministic user-defined language model on top of (or instead of) the N-gram, registered the way
   julius/main.c:153-161 does (JREF_USERLM=1 together with the -userlm option) */
static LOGPROB my_uni(WORD_INFO *winfo, WORD_ID w, LOGPROB ngram_prob) {
  return ngram_prob * 0.8f - 0.01f * (float)(w % 13);
}
static LOGPROB my_bi(WORD_INFO *winfo, WORD_ID context, WORD_ID w, LOGPROB ngram_prob) {
  return ngram_prob * 0.9f - 0.02f * (float)(((int)context * 7 + (int)w * 3) % 11);
}
static LOGPROB my_lm(WORD_INFO *winfo, WORD_ID *contexts, int context_len, WORD_ID w, LOGPROB ngram_prob) {
  return ngram_prob;
}

/* progressive output (-progout): what bt_current_max left in r->result.pass1 (beam.c:876-921), every interval */
static void on_interim(Recog *recog, void *dummy) {
  RecogProcess *r = recog->process_list;
  union { float f; unsigned u; } sc;
  int i;
  if (!r->have_interim) return;
  sc.f = (r->result.pass1.word_num > 0) ? r->result.pass1.score : 0.0f;
  fprintf(stdout, "JREF_INTERIM utt=%d frame=%d score=%08x words=", g_utt, r->result.num_frame, sc.u);
  for (i = 0; i < r->result.pass1.word_num; i++) fprintf(stdout, "%s%d", i ? "," : "", (int)r->result.pass1.word[i]);
  fprintf(stdout, "\n");
}

int main(int argc, char *argv[]) {
  Jconf *jconf;
  Recog *recog;
  int ret, i, nargs = 0;
  char **args;
  const char *dump = NULL;
  char fname[MAXPATHLEN];

  args = (char **)malloc(sizeof(char *) * (argc + 1));
  for (i = 0; i < argc; i++) {
    if (i > 0 && strcmp(argv[i], "-dump") == 0 && i + 1 < argc) { dump = argv[++i]; continue; }
    args[nargs++] = argv[i];
  }
  if (dump == NULL) { fprintf(stderr, "usage: jref -dump out.jrf [julius options]\n"); return 2; }
  signal(SIGSEGV, on_crash); signal(SIGBUS, on_crash); signal(SIGABRT, on_crash);
  if (getenv("JREF_TOKENS")) g_tokens = atoi(getenv("JREF_TOKENS"));
  if (getenv("JREF_QUIET")) jlog_set_output(NULL);
  g_out = fopen(dump, "wb");
  if (!g_out) { perror(dump); return 2; }

  jconf = j_config_load_args_new(nargs, args);
  if (jconf == NULL) return 1;
  if (getenv("JREF_USERLM")) {
    /* j_create_instance_from_jconf in its three steps, with the user LM registered between loading and fusion */
    PROCESS_LM *lm;
    if (j_jconf_finalize(jconf) == FALSE) return 1;
    recog = j_recog_new();
    recog->jconf = jconf;
    if (j_load_all(recog, jconf) == FALSE) { fprintf(stderr, "jref: error in loading model\n"); return 1; }
    for (lm = recog->lmlist; lm; lm = lm->next) if (lm->lmtype == LM_PROB) j_regist_user_lm_func(lm, my_uni, my_bi, my_lm);
    if (j_final_fusion(recog) == FALSE) { fprintf(stderr, "jref: error in startup\n"); return 1; }
  } else
  recog = j_create_instance_from_jconf(jconf);
  if (recog == NULL) { fprintf(stderr, "jref: error in startup\n"); return 1; }
  callback_add(recog, CALLBACK_EVENT_PASS1_BEGIN, on_pass1_begin, NULL);
  callback_add(recog, CALLBACK_EVENT_PASS1_FRAME, on_pass1_frame, NULL);
  callback_add(recog, CALLBACK_EVENT_PASS1_END, on_pass1_end, NULL);
  if (getenv("JREF_RESULT")) callback_add(recog, CALLBACK_RESULT, on_result, NULL);
  if (getenv("JREF_INTERIM")) callback_add(recog, CALLBACK_RESULT_PASS1_INTERIM, on_interim, NULL);
  if (j_adin_init(recog) == FALSE) return 1;

  if (jconf->input.speech_input == SP_MFCFILE || jconf->input.speech_input == SP_OUTPROBFILE) {
    /* file names come from stdin, one per line (julius-simple.c:292-309) */
    while (fgets(fname, MAXPATHLEN, stdin) != NULL) {
      char *p = fname + strlen(fname);
      while (p > fname && (p[-1] == '\n' || p[-1] == '\r' || p[-1] == ' ')) *--p = '\0';
      if (fname[0] == '\0') continue;
      ret = j_open_stream(recog, fname);
      if (ret == -1) continue;
      if (ret == -2) break;
      ret = j_recognize_stream(recog);
      if (ret == -1) return 1;
    }
  } else {
    fprintf(stderr, "jref: only -input mfcfile / outprob is supported by this driver\n");
    return 2;
  }
  fprintf(stdout, "JREF_SUMMARY utts=%d frames=%ld decode_sec=%.6f\n", g_utt, g_frames, g_decode_sec);
  fclose(g_out);
  return 0;
}
