/* oracle/restate/gmm.c -- TEST INFRASTRUCTURE (CPU restatement, see oracle.h).
 *
 * State-level GMM log-likelihood exactly as the reference computes it:
 *   compute_g_base   libsent/src/phmm/gprune_none.c:58-82
 *   gprune_none      libsent/src/phmm/gprune_none.c:132-182
 *   compute_g_safe   libsent/src/phmm/gprune_safe.c:75-95
 *   gprune_safe      libsent/src/phmm/gprune_safe.c:159-202 (last_id == NULL branch;
 *                    gprune_beam.c:340-354 and gprune_heu.c:327-341 are the same code)
 *   cache_push       libsent/src/phmm/gprune_common.c:41-126
 *   calc_mix         libsent/src/phmm/calc_mix.c:40-81
 *   addlog_array     libsent/src/phmm/addlog.c:28-57,102-123
 *   outprob_cd_*     libsent/src/phmm/outprob.c:286-400
 * All arithmetic is fp32 with the same double promotions the C source implies
 * (compile with -ffp-contract=off, no -ffast-math).
 */
#include <math.h>
#include "oracle.h"

#define TBLSIZE 500000
#define VRANGE 15
#define TMAG 33333.3333

static float g_tbl[TBLSIZE];
static int g_tbl_built = 0;

void oracle_addlog_table(float *out) {
  int i;
  if (!g_tbl_built) {
    for (i = 0; i < TBLSIZE; i++) {
      float f = -((float)VRANGE * (float)i / (float)TBLSIZE);
      g_tbl[i] = log(1 + exp(f));
    }
    g_tbl_built = 1;
  }
  if (out) memcpy(out, g_tbl, sizeof(g_tbl));
}

float oracle_addlog_array(const float *a, int n) {
  float tmp, x, y;
  unsigned int idx;
  oracle_addlog_table(NULL);
  y = JB200_LOG_ZERO;
  for (n--; n >= 0; n--) {
    x = a[n];
    if (x > y) { tmp = x; x = y; y = tmp; }
    if ((tmp = x - y) < JB200_LOG_ADDMIN) continue;
    else {
      idx = (unsigned int)((-tmp) * TMAG + 0.5);
      y += g_tbl[idx];
    }
  }
  return y;
}

static float g_base(const float *vec, const float *mean, const float *var, float gconst, int veclen) {
  float tmp, x;
  tmp = gconst;
  for (; veclen > 0; veclen--) {
    x = *(vec++) - *(mean++);
    tmp += x * x * *(var++);
  }
  return (tmp * -0.5);
}

static float g_safe(const float *vec, const float *mean, const float *var, float gconst, int veclen, float thres) {
  float tmp, x;
  float fthres = thres * (-2.0);
  tmp = gconst;
  for (; veclen > 0; veclen--) {
    x = *(vec++) - *(mean++);
    tmp += x * x * *(var++);
    if (tmp > fthres) return JB200_LOG_ZERO;
  }
  return (tmp * -0.5);
}

static int find_insert_point(const float *calced_score, float score, int len) {
  int left = 0, right = len - 1, mid;
  while (left < right) {
    mid = (left + right) / 2;
    if (calced_score[mid] > score) left = mid + 1; else right = mid;
  }
  return left;
}

static int cache_push(float *calced_score, int *calced_id, int gprune_num, int id, float score, int len) {
  int insertp;
  if (len == 0) { calced_score[0] = score; calced_id[0] = id; return 1; }
  if (calced_score[len - 1] >= score) {
    if (len < gprune_num) { calced_score[len] = score; calced_id[len] = id; len++; }
    return len;
  }
  if (calced_score[0] < score) insertp = 0;
  else insertp = find_insert_point(calced_score, score, len);
  if (len < gprune_num) {
    memmove(&calced_score[insertp + 1], &calced_score[insertp], sizeof(float) * (len - insertp));
    memmove(&calced_id[insertp + 1], &calced_id[insertp], sizeof(int) * (len - insertp));
  } else if (insertp < len - 1) {
    memmove(&calced_score[insertp + 1], &calced_score[insertp], sizeof(float) * (len - insertp - 1));
    memmove(&calced_id[insertp + 1], &calced_id[insertp], sizeof(int) * (len - insertp - 1));
  }
  calced_score[insertp] = score;
  calced_id[insertp] = id;
  if (len < gprune_num) len++;
  return len;
}

int oracle_gmm_score(const jb200_gmm_desc *g, const float *feat, int T, float *out) {
  int S = g->n_states, D = g->dim, t, s, i;
  float *score = (float *)malloc(sizeof(float) * (g->max_mix + 1));
  int *id = (int *)malloc(sizeof(int) * (g->max_mix + 1));
  oracle_addlog_table(NULL);
  for (t = 0; t < T; t++) {
    const float *vec = feat + (size_t)t * D;
    for (s = 0; s < S; s++) {
      int g0 = g->state_off[s], gnum = g->state_off[s + 1] - g0, num = 0;
      float logprob, logprobsum = 0.0;
      if (g->gprune_method == JB200_GPRUNE_NONE) {
        for (i = 0; i < gnum; i++) {
          int k = g0 + i;
          score[i] = g->valid[k] ? g_base(vec, g->mean + (size_t)k * D, g->ivar + (size_t)k * D, g->gconst[k], D) : JB200_LOG_ZERO;
          id[i] = i;
        }
        num = gnum;
      } else {
        float thres = JB200_LOG_ZERO, sc;
        for (i = 0; i < gnum; i++) {
          int k = g0 + i;
          if (num < g->gprune_num) {
            sc = g->valid[k] ? g_base(vec, g->mean + (size_t)k * D, g->ivar + (size_t)k * D, g->gconst[k], D) : JB200_LOG_ZERO;
          } else {
            sc = g->valid[k] ? g_safe(vec, g->mean + (size_t)k * D, g->ivar + (size_t)k * D, g->gconst[k], D, thres) : JB200_LOG_ZERO;
            if (sc <= thres) continue;
          }
          num = cache_push(score, id, g->gprune_num, i, sc, num);
          thres = score[num - 1];
        }
      }
      for (i = 0; i < num; i++) score[i] += g->lnweight[g0 + id[i]];
      logprob = oracle_addlog_array(score, num);
      if (!(logprob <= JB200_LOG_ZERO)) logprobsum += logprob * 1.0f;
      if (logprobsum == 0.0) out[(size_t)t * S + s] = JB200_LOG_ZERO;
      else if (logprobsum <= JB200_LOG_ZERO) out[(size_t)t * S + s] = JB200_LOG_ZERO;
      else out[(size_t)t * S + s] = (logprobsum * JB200_INV_LOG_TEN);
    }
  }
  free(score); free(id);
  return 0;
}

/* one pseudo-phone set (outprob.c:286-400); exported for the beam restatement */
float oracle_cdset_one(const jb200_gmm_desc *g, const float *strow, int c, float *nbest_work) {
  int b0 = g->cd_off[c], n_in = g->cd_off[c + 1] - b0, i, k, n;
  float prob;
  switch (g->iwcd_method) {
    case JB200_IWCD_AVG: {
      float sum = 0.0; int j = 0;
      for (i = 0; i < n_in; i++) { float p = strow[g->cd_states[b0 + i]]; if (p > JB200_LOG_ZERO) { sum += p; j++; } }
      return sum / (float)j;
    }
    case JB200_IWCD_MAX: {
      float maxprob = JB200_LOG_ZERO;
      for (i = 0; i < n_in; i++) { prob = strow[g->cd_states[b0 + i]]; if (maxprob < prob) maxprob = prob; }
      return maxprob;
    }
    default: {
      int maxn = g->iwcd_nbest;
      float *mp = nbest_work;
      n = 0;
      for (i = 0; i < n_in; i++) {
        prob = strow[g->cd_states[b0 + i]];
        if (prob <= JB200_LOG_ZERO) continue;
        if (n == 0 || prob <= mp[n - 1]) {
          if (n == maxn) continue;
          mp[n] = prob; n++;
        } else {
          for (k = 0; k < n; k++) {
            if (prob > mp[k]) {
              memmove(&mp[k + 1], &mp[k], sizeof(float) * (n - k - ((n == maxn) ? 1 : 0)));
              mp[k] = prob;
              break;
            }
          }
          if (n < maxn) n++;
        }
      }
      prob = 0.0;
      for (i = 0; i < n; i++) prob += mp[i];
      return prob / (float)n;
    }
  }
}

int oracle_cdset_score(const jb200_gmm_desc *g, const float *st, int T, float *out) {
  int S = g->n_states, C = g->n_cdsets, t, c;
  float *work = (float *)malloc(sizeof(float) * (g->iwcd_nbest + 2));
  for (t = 0; t < T; t++)
    for (c = 0; c < C; c++)
      out[(size_t)t * C + c] = oracle_cdset_one(g, st + (size_t)t * S, c, work);
  free(work);
  return 0;
}
