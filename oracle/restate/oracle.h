/* oracle/restate/oracle.h -- TEST INFRASTRUCTURE.  Not part of the product.
 *
 * CPU restatement of the reference's hot path on the flattened arrays of
 * include/jb200_model.h.  Plain sequential C that follows the reference's
 * arithmetic statement by statement (every function cites the file:line it
 * restates) so that results are BIT-IDENTICAL to the compiled reference
 * (oracle/_ref) -- that identity is what tests/test_oracle_*.py pin.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product never does.
 */
#ifndef JB200_ORACLE_H
#define JB200_ORACLE_H
#include "jb200_model.h"
#ifdef __cplusplus
extern "C" {
#endif

/* trellis atom as the parity harness sees it (libjulius/include/julius/trellis.h:28-45) */
typedef struct {
  int32_t wid, begintime, endtime;
  float backscore, lscore;
  int32_t last;          /* index of last_tre in the same array, -1 = sentence start */
} oracle_atom;

void oracle_addlog_table(float *tbl500k);                 /* addlog.c:39-57 */
float oracle_addlog_array(const float *a, int n);         /* addlog.c:102-123 */

/* GMM state scores: feat [T][D] -> out [T][S] log10 (calc_mix.c:40-81 over gprune_none/safe) */
int oracle_gmm_score(const jb200_gmm_desc *g, const float *feat, int T, float *out);
/* pseudo-phone set scores from state scores: st [T][S] -> out [T][C] (outprob.c:286-400) */
int oracle_cdset_score(const jb200_gmm_desc *g, const float *st, int T, float *out);
/* DNN forward: in [T][in_dim] -> out [T][out_dim] log10 pseudo-likelihood (calc_dnn.c:774-868) */
int oracle_dnn_score(const jb200_dnn_desc *d, const float *in, int T, float *out);

/* pass-1 beam over one utterance given the state-score matrix st [T][S]
 * (beam.c:1825-3162).  Returns number of atoms written (<= max_atoms), or <0.
 * atoms are emitted in the finalized order (frame-major, wid-sorted within a frame).
 * best_words (<=150, reverse order as the reference stores them), *n_best, *best_score,
 * *status (0 ok, -1 search failed). */
int oracle_beam_decode(const jb200_tree_desc *t, const jb200_gmm_desc *g,
                       const float *st, int T, int S,
                       oracle_atom *atoms, int max_atoms,
                       int *best_words, int *n_best, float *best_score, int *status,
                       /* optional per-frame survivor trace, may be NULL */
                       int *trace_counts /*[T][2]: tnum, nsurv*/);
#ifdef __cplusplus
}
#endif
#endif
