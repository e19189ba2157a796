// beam.cu -- K3: pass-1 lexicon-tree token passing, one persistent thread block per utterance.
//
// Stands in for libjulius/src/beam.c get_back_trellis_init/_proceed/_end + finalize_1st_pass
// (:1825, :2663, :3052, :3133), outprob_style (outprob_style.c:354-494), the factoring look-ups
// (factoring_sub.c:942-1143, ngram_access.c:249-305) and the word-trellis store/sort
// (backtrellis.c:190-267,438-478), stock "fast" switches: N-gram LMs on normal trees (beam_kernel, body in
// beam_frames.inc) and multipath trees (beam_kernel_mp); DFA grammars on the category tree (beam_kernel_grammar = the
// same body with the three grammar-mode differences compiled in).
//
// Why this is not a transliteration.  The reference walks the survivors of frame t-1 one by one
// and lets each arc "propagate" into a per-node slot; ties are won by whoever arrived first, new
// tokens are numbered in order of first arrival, and the beam is cut by an in-place heap select
// whose OUTPUT ORDER (it decides next frame's arrival order) depends on the heap's mechanics.
// Scores sit on a coarse fp32 grid (|score| ~ 3e4 => ulp 2^-9), so exact ties are routine and
// every one of these order effects is observable in the word trellis.  The kernel computes the
// same fixed point with parallel primitives that are order-equivalent by construction:
//   * every candidate transition gets the sequence number it would have had in the sequential
//     walk:  seq = (survivor position j) * 2^18 + (arc index | n_intra + isolated-root index);
//     the factoring pass gets j = n_survivors;
//   * per destination node:  first arrival   = atomicMin(seq)          (who creates the token)
//                            winning content = atomicMax(score, ~seq)  (strict '<' keeps the incumbent)
//   * new-token numbering = rank of the creators in arrival order (one bit per candidate position, popcount prefix);
//   * inter-word transitions into the "isolated" roots are pre-reduced per root over the frame's
//     word-end tokens (max is associative; first/winner seqs are carried along); the per-last-word
//     bigram rows the reference caches lazily (iw_sc_cache) are tabulated once at create time;
//   * the beam cut replays the reference's heap select exactly: bottom-up heap construction is
//     level-parallel (siftdowns of one level touch disjoint subtrees), the extractions are
//     replayed by one thread on a shared-memory heap (heap_extract_fast) while the other warps
//     reset the per-node slots for the next frame.
// All frames of an utterance run inside one kernel launch; tokens, node slots and candidates
// live in per-utterance global scratch (L2 resident), the sort/heap array in shared memory.
//
// Compiled with --fmad=false: every float decision uses the reference's fp32 expression order.
#include "common.cuh"
#include "heap_pipe.cuh"
#include <vector>
#include <algorithm>

struct jb200_gmm;
struct jb200_dnn;
namespace jb200 {
int dnn_forward_device(jb200_dnn *h, const float *d_in, int T, float *d_rows, int row_stride, cudaStream_t st);
int gmm_device(const jb200_gmm *h);
int gmm_dim(const jb200_gmm *h);
int gmm_launch_states(jb200_gmm *h, const float *d_feats, int T, float *d_rows, int row_stride, cudaStream_t st,
                      const int *seg_off, const int *seg_start, int n_seg);
int gmm_cd_device(const jb200_gmm *h, const int **cd_off, const int **cd_states, int *method, int *nbest);

#ifndef JB200_BEAM_THREADS
#define JB200_BEAM_THREADS 256
#endif
static constexpr int BEAM_THREADS = JB200_BEAM_THREADS;
static constexpr int NWARP = BEAM_THREADS / 32;
static constexpr int SEQ_LOCAL_BITS = 18;
static constexpr unsigned SEQ_LOCAL = 1u << SEQ_LOCAL_BITS;
static constexpr int CD_NMAX = 16;
static constexpr int MAX_WORDS = 150;      // MAXSEQNUM, libsent/include/sent/speech.h:50

struct __align__(16) NodeRec { float self_a, next_a; int arc_off, arc_n; int stend, next; int scid, out; };   // 32 B; (scid,out) is one aligned 8-byte word; next = the node next_a leads to
struct __align__(8) Tok { float score; int node; int tre; int cword; float lscore; int tre_wid; };              // 24 B
struct __align__(16) Cand { float score; int node; float lscore; int src; };                                     // 16 B
struct __align__(16) CandB { int tre; int cword; int tre_wid; int out; };                                        // 16 B: what the winner hands to the new token
struct __align__(16) IsoCand { float score; int e; float lscore; int first_e; };                                 // 16 B
struct __align__(8) WEnd { int j; int atom; int last_word; float base; int transp2; int nintra; };              // 24 B

// per-node arrival slot: who reached the node first (creator) and who reached it best (content); one aligned
// 16-byte record so that the two atomics, the reset and the later reads of a node touch a single sector
struct __align__(16) NodeSlot { unsigned long long bestkey; int firstseq; int pad; };
struct SlotView {
  NodeSlot *s;
  __device__ __forceinline__ int *fs(int node) const { return &s[node].firstseq; }
  __device__ __forceinline__ unsigned long long *bk(int node) const { return &s[node].bestkey; }
  __device__ __forceinline__ void set(int node, int firstseq, unsigned long long bestkey) const {
    __stcg(reinterpret_cast<uint4 *>(s + node), make_uint4((unsigned)bestkey, (unsigned)(bestkey >> 32), (unsigned)firstseq, 0u));
  }
  __device__ __forceinline__ void reset(int node) const { set(node, 0x7fffffff, 0ull); }
};

// One launch of a beam kernel covers frames [t0, t1) of an utterance: the whole utterance (FIRST|FINAL), or one piece of
// it -- the batch pipeline cuts utterances into chunks so that the scoring of chunk c+1 runs beside the token passing of
// chunk c, and a stream (jb200_stream_*) advances as its input arrives, which is how the reference drives pass 1
// (decode_proceed, one frame per call, libjulius/src/pass1.c:112-254).  Everything an utterance carries from frame to
// frame lives in its global work area already; the few scalars the kernel keeps in shared memory are parked in UttState.
static constexpr int CHUNK_FIRST = 1, CHUNK_FINAL = 2, CHUNK_SKIP = 4;   // SKIP: nothing to do for this utterance in this launch
struct ChunkDesc { int t0, t1, flags, row_base; };   // score row of frame t: rows + (row_base + t) * row_stride
struct UttState {
  int ns, natoms, tnum_prev, slots_clean, overflow, stopped, cur, n_left;
  float thr; int t_done;
  // best partial sentence at the last frame done (bt_current_max, beam.c:876-921): filled when BeamParams.interim is set
  int interim_frame, interim_nwords; float interim_score; int pad_;
  long long prof[8];
};

struct BeamParams {
  // tree
  const NodeRec *nodes; const int *arc_to; const float *arc_a;
  const int *rset_ctx; const int *word_ctx; int n_ctx;
  const int *iso_node; const int *iso_id; int n_iso; const float *iw;
  const int *shared_node; const float *shared_f; int n_shared;
  // multipath trees: the roots carry no output, so cross-word transitions land one arc further
  // (beam.c:2467-2500, :2584-2605): the successors of the isolated / shared roots, root-major
  int multipath; const int *isoarc_node; const int *isoarc_iso; const float *isoarc_a; int n_isoarc;
  const int *sharc_node; const int *sharc_shared; const float *sharc_a; int n_sharc;
  const float *wordend_a; const uint8_t *is_transp; const int *wton; const float *cprob;
  const float *fscore; const int *scword;
  const float *uni_prob; const float *uni_bow; const int *bi_bgn; const int *bi_num; const int *bi_wid; const float *bi_prob;
  int lm_mode, lm_unk_id; float lm_unk_num_log;
  float lm_weight, lm_penalty, lm_penalty_trans, prune_width;
  int head_node, tail_silwid, beam, n_nodes;
  // cd sets
  const int *cd_off; const int *cd_states; int iwcd_method, iwcd_nbest;
  // batch
  const float *rows; int row_stride; const int *frame_off;
  // per-utterance work areas (index = blockIdx.x)
  Tok *tok; int *order; NodeSlot *slots; Cand *cand; CandB *candb; Tok *surv; IsoCand *iso; WEnd *wend;
  jb200_atom *atoms_raw; int *newidx; int *group0; int *counts;
  const long long *atom_off;
  // compact outputs
  jb200_atom *atoms_out; unsigned long long *atom_counter; long long atoms_out_cap;
  jb200_utt_result *results; int *words;
  long long *prof;            // [n_utts][8] cycle counters per phase, or NULL
  unsigned *bitmask; int *wordpre;   // per-utterance arrival-order bitmask [maxbits/32] and its word prefix counts
  unsigned long long *misspec_counter; int force_seq_heap, check_heap, no_lose, prof_fine, heap_single, no_closed;
  unsigned long long *lmc; int lmc_bits;      // memo of max_successor_prob, 2^lmc_bits entries (0 = off)
  int maxt, maxc, maxw, maxbits;
  // token sets too large for shared memory (wide beams on large trees): the heap-select array lives in global memory
  // ([n_utts][maxt+4] entries) and shared memory only holds the closed form's sort area (sort_cap 8-byte keys)
  unsigned long long *heap_g; int sort_cap, qcap;      // qcap: 8-byte entries of the shared-memory area in front of offs (>= sort_cap)
  // grammar (DFA) mode, appended so that the offsets of everything above stay what the N-gram kernels were built with
  const uint8_t *cp_allowed; const int *init_node; const float *init_lscore; int n_init; float penalty1;
  // chunked launches
  const ChunkDesc *chunk; UttState *state; int interim; int *interim_words;   // [n_utts][MAX_WORDS]
  int no_reloc;             // JB200_NO_RELOCATE=1: replay the loop whenever the plain closed form's test fails (A/B)
  int atoms_in_place;       // streams: the finalized atoms of utterance u go to atoms_out + atom_off[u] (no batch compaction)
};

// ---- small device helpers ----------------------------------------------------------------------
__device__ __forceinline__ unsigned fkey(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ int search_bigram(const BeamParams &p, int wc, int w) {
  int left = __ldg(p.bi_bgn + wc);
  if (left < 0) return -1;
  int right = left + __ldg(p.bi_num + wc) - 1;
  while (left < right) {
    int mid = (left + right) / 2;
    if (__ldg(p.bi_wid + mid) < w) left = mid + 1; else right = mid;
  }
  return (__ldg(p.bi_wid + left) == w) ? left : -1;
}

// bi_prob_normal / _additional(_oldbin) / _compute, ngram_access.c:288-405
__device__ float bigram_prob(const BeamParams &p, int w1, int w2) {
  int n2; float prob;
  if (p.lm_mode == JB200_BI_NORMAL || p.lm_mode == JB200_BI_ADDITIONAL_OLDBIN) {
    if ((n2 = search_bigram(p, w1, w2)) >= 0) prob = __ldg(p.bi_prob + n2);
    else prob = __ldg(p.uni_bow + w1) + __ldg(p.uni_prob + w2);
  } else if (p.lm_mode == JB200_BI_ADDITIONAL) {
    if ((n2 = search_bigram(p, w2, w1)) >= 0) prob = __ldg(p.bi_prob + n2);
    else prob = __ldg(p.uni_bow + w1) + __ldg(p.uni_prob + w2);
  } else {
    if ((n2 = search_bigram(p, w2, w1)) >= 0) prob = __ldg(p.bi_prob + n2);
    else prob = __ldg(p.uni_bow + w2) + __ldg(p.uni_prob + w1);
    prob = prob + __ldg(p.uni_prob + w2) - __ldg(p.uni_prob + w1);
  }
  if (w2 != p.lm_unk_id) return prob;
  return prob - p.lm_unk_num_log;
}

// max_successor_prob, factoring_sub.c:942-1012 (1-gram factoring build).  The reference keeps a per-node
// one-entry cache of the bigram look-up (factoring_sub.c:984-1006); here a direct-mapped table shared by
// all utterances memoises the pure function (last word, successor-word slot) -> value, because the binary
// search is ~10 dependent global loads and the same pairs recur every frame while a token waits on the
// node in front of a single-word branch.  One 64-bit word per entry: key in the high half, value bits in
// the low half, written and read atomically as a unit.
__device__ __forceinline__ float max_successor_prob(const BeamParams &p, int lastword, int scid) {
  if (lastword < 0) return 0.0f;
  if (scid < 0) return __ldg(p.fscore - scid);
  unsigned long long *slot = nullptr;
  unsigned key = 0u;
  if (p.lmc_bits > 0) {
    key = ((unsigned)lastword << 16) | (unsigned)scid;
    slot = p.lmc + ((key * 2654435761u) >> (32 - p.lmc_bits));
    const unsigned long long e = __ldcg(slot);
    if ((unsigned)(e >> 32) == key) return __uint_as_float((unsigned)e);
  }
  int w = __ldg(p.scword + scid);
  const float v = bigram_prob(p, __ldg(p.wton + lastword), __ldg(p.wton + w)) + __ldg(p.cprob + w);
  if (slot) __stcg(slot, ((unsigned long long)key << 32) | __float_as_uint(v));
  return v;
}

// a state score is read once per utterance-frame (the row belongs to this utterance alone): JB200_STREAM_HINTS marks
// these and the last reads of the candidate records evict-first, so that they displace less of what is re-used
#ifdef JB200_STREAM_HINTS
#define JB_LD_ROW(p_) __ldcs(p_)
#else
#define JB_LD_ROW(p_) __ldg(p_)
#endif
// outprob_cd, outprob.c:286-400, evaluated on demand from the frame's state-score row
__device__ float cdset_score(const BeamParams &p, const float *__restrict__ row, int c) {
  const int b0 = __ldg(p.cd_off + c), n_in = __ldg(p.cd_off + c + 1) - b0;
  if (p.iwcd_method == JB200_IWCD_AVG) {
    float sum = 0.0f; int j = 0;
    for (int i = 0; i < n_in; i++) { float v = JB_LD_ROW(row + __ldg(p.cd_states + b0 + i)); if (v > JB200_LOG_ZERO) { sum += v; j++; } }
    return sum / (float)j;
  } else if (p.iwcd_method == JB200_IWCD_MAX) {
    float mx = JB200_LOG_ZERO;
    for (int i = 0; i < n_in; i++) { float v = JB_LD_ROW(row + __ldg(p.cd_states + b0 + i)); if (mx < v) mx = v; }
    return mx;
  }
  const int maxn = p.iwcd_nbest;
  if (maxn <= 3) {
    // the kept list is the sorted multiset of the maxn largest valid scores, and the result adds it up from the
    // largest down (outprob.c:313-318): three registers instead of an indexed array in local memory
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY; int n = 0;
    for (int i = 0; i < n_in; i++) {
      const float v = JB_LD_ROW(row + __ldg(p.cd_states + b0 + i));
      if (v <= JB200_LOG_ZERO) continue;
      n++;
      if (v > m0) { m2 = m1; m1 = m0; m0 = v; }
      else if (v > m1) { m2 = m1; m1 = v; }
      else if (v > m2) m2 = v;
    }
    n = min(n, maxn);
    float prob = 0.0f;
    if (n > 0) prob += m0;
    if (n > 1) prob += m1;
    if (n > 2) prob += m2;
    return prob / (float)n;
  }
  float mp[CD_NMAX + 1]; int n = 0;
  for (int i = 0; i < n_in; i++) {
    float prob = JB_LD_ROW(row + __ldg(p.cd_states + b0 + i));
    if (prob <= JB200_LOG_ZERO) continue;
    if (n == 0 || prob <= mp[n - 1]) {
      if (n == maxn) continue;
      mp[n] = prob; n++;
    } else {
      for (int k = 0; k < n; k++) {
        if (prob > mp[k]) {
          int cnt = n - k - ((n == maxn) ? 1 : 0);
          for (int q = k + cnt; q > k; q--) mp[q] = mp[q - 1];
          mp[k] = prob;
          break;
        }
      }
      if (n < maxn) n++;
    }
  }
  float prob = 0.0f;
  for (int i = 0; i < n; i++) prob += mp[i];
  return prob / (float)n;
}

// outprob_style, outprob_style.c:354-494 with the context resolution tabulated on the host
__device__ __forceinline__ float outprob_style(const BeamParams &p, const float *__restrict__ row, int out, int last_wid) {
  const int style = (unsigned)out >> 28, ref = out & 0x0fffffff;
  if (style == JB200_AS_STATE) return JB_LD_ROW(row + ref);
  if (style == JB200_AS_LSET) return cdset_score(p, row, ref);
  const int col = (last_wid < 0) ? p.n_ctx : __ldg(p.word_ctx + last_wid);
  const int r = __ldg(p.rset_ctx + (size_t)ref * (p.n_ctx + 1) + col);
  if (r >= 0) return JB_LD_ROW(row + r);
  return cdset_score(p, row, -r - 1);
}

// block-wide exclusive scan of one int per thread; *total = block sum.  Ends with a barrier.
__device__ __forceinline__ int block_excl_scan(int v, int *warp_sums, int *total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int s = (lane < NWARP) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    if (lane < NWARP) warp_sums[lane] = s;   // inclusive
  }
  __syncthreads();
  const int base = (wid == 0) ? 0 : warp_sums[wid - 1];
  *total = warp_sums[NWARP - 1];
  const int r = base + x - v;
  __syncthreads();
  return r;
}

__device__ __forceinline__ void cand_atomics(const SlotView &slots,
                                             int node, float score, unsigned seq_first, unsigned seq_win) {
  atomicMin(slots.fs(node), (int)seq_first);
  unsigned long long key = ((unsigned long long)fkey(score) << 32) | (unsigned)(~seq_win);
  atomicMax(slots.bk(node), key);
}

// heap entries: high 32 bits = token id, low 32 bits = fp32 score bits.  Heap index h (1-based) lives
// in slot h, so the sibling pair (2p, 2p+1) is one aligned 16-byte word.
__device__ __forceinline__ float hval(unsigned long long e) { return __uint_as_float((unsigned)(e & 0xffffffffu)); }

template <bool MAXHEAP>
__device__ __forceinline__ bool hcmp(float a, float b) { return MAXHEAP ? (a < b) : (a > b); }      // "child < child+1"
template <bool MAXHEAP>
__device__ __forceinline__ bool hstop(float s, float c) { return MAXHEAP ? (s >= c) : (s <= c); }   // "STVAL >= SVAL(child)"

template <bool MAXHEAP>
__device__ __forceinline__ void sift_down(unsigned long long *A, int start, int n) {
  // the inner loop of sort_token_upward / _downward, beam.c:1355-1368 / :1421-1434
  const unsigned long long s = A[start];
  const float sv = hval(s);
  int parent = start, child;
  while ((child = parent * 2) <= n) {
    const ulonglong2 pr = *reinterpret_cast<const ulonglong2 *>(A + child);
    unsigned long long c = pr.x;
    if (child < n && hcmp<MAXHEAP>(hval(pr.x), hval(pr.y))) { child++; c = pr.y; }
    if (hstop<MAXHEAP>(sv, hval(c))) break;
    A[parent] = c;
    parent = child;
  }
  A[parent] = s;
}

// one extraction step chain, two tree levels per shared-memory round trip
template <bool MAXHEAP>
__device__ __forceinline__ void sift_root_2level(unsigned long long *A, const unsigned long long s, const int m) {
  const float sv = hval(s);
  int parent = 1;
  while (true) {
    int child = parent * 2;
    if (child > m) break;
    // children pair and the four grandchildren (slots 4p..4p+3), loaded together
    const ulonglong2 pr = *reinterpret_cast<const ulonglong2 *>(A + child);
    const int g = parent * 4;
    ulonglong2 g01 = make_ulonglong2(0ull, 0ull), g23 = make_ulonglong2(0ull, 0ull);
    if (g <= m) g01 = *reinterpret_cast<const ulonglong2 *>(A + g);
    if (g + 2 <= m) g23 = *reinterpret_cast<const ulonglong2 *>(A + g + 2);
    // level 1
    unsigned long long c = pr.x; bool right = false;
    if (child < m && hcmp<MAXHEAP>(hval(pr.x), hval(pr.y))) { right = true; c = pr.y; }
    if (hstop<MAXHEAP>(sv, hval(c))) break;
    A[parent] = c;
    parent = child + (right ? 1 : 0);
    // level 2 (children of the chosen child are g01 or g23)
    child = parent * 2;
    if (child > m) break;
    const ulonglong2 q = right ? g23 : g01;
    unsigned long long c2 = q.x; bool right2 = false;
    if (child < m && hcmp<MAXHEAP>(hval(q.x), hval(q.y))) { right2 = true; c2 = q.y; }
    if (hstop<MAXHEAP>(sv, hval(c2))) break;
    A[parent] = c2;
    parent = child + (right2 ? 1 : 0);
  }
  A[parent] = s;
}

template <bool MAXHEAP>
__device__ void heap_build(unsigned long long *A, int n) {
  // build: roots n/2 .. 1; all roots of one tree level are independent (disjoint subtrees) and the
  // sequential order visits deeper levels first, so a level-synchronous sweep is equivalent.
  const int half = n >> 1;
  if (half >= 1) {
    for (int L = 31 - __clz(half); L >= 0; L--) {
      const int lo = 1 << L;
      const int hi = min((2 << L) - 1, half);
      for (int i = lo + (int)threadIdx.x; i <= hi; i += BEAM_THREADS) sift_down<MAXHEAP>(A, i, n);
      __syncthreads();
    }
  }
}

// Sequential extraction replay (beam.c:1370-1384): s = last; last = root; shrink; sift s from the root.
template <bool MAXHEAP>
__device__ void heap_extract_seq(unsigned long long *A, int n, int extract) {
  if (threadIdx.x == 0) {
    int m = n;
    while (m > n - extract) {
      const unsigned long long s = A[m];
      A[m] = A[1];
      m--;
      if (m >= 1) sift_root_2level<MAXHEAP>(A, s, m);
    }
  }
  __syncthreads();
}

// The pipelined replay (heap_pipe.cuh) for a heap in GLOBAL memory -- the token set of a wide beam on a large tree does
// not fit shared memory (-b 4000 on the 60k-word tree: up to 34k tokens a frame).  Same schedule, plain generic loads and
// stores that bypass L1 (another lane wrote the line one tick ago); a tick costs an L2 round trip instead of a
// shared-memory one, so this path is for the frames the closed form cannot answer.  outs: shared memory.
__device__ __forceinline__ ulonglong2 ldcg_pair(const unsigned long long *p) {
  const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p));
  return make_ulonglong2(((unsigned long long)v.y << 32) | v.x, ((unsigned long long)v.w << 32) | v.z);
}
// Shared-memory copies that go with a global-memory heap (write-through, so global stays the truth):
//   top  : slots [0, cs) -- the top levels of the tree, which every extraction walks;
//   tail : slots [tail_first, tail_first + tail_n) -- where the extractions take their s from.
// Deeper levels are touched only by the few sifts that follow winners all the way down.
struct HeapCache {
  unsigned long long *top; int cs;
  unsigned long long *tail; int tail_first, tail_n;
};

template <bool MAXHEAP>
__device__ void heap_extract_pipe_global(unsigned long long *A, const int n, const int extract, const float lose_below,
                                         unsigned long long *outs, const int maxt, const HeapCache hc,
                                         unsigned &ticks_out, unsigned &stalls_out) {
  constexpr int NL = 16;
  constexpr unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long sent = MAXHEAP ? 0xff800000ull : 0x7f800000ull;
  const int cap = (maxt >> 1) + 1;                          // pair (maxt+2, maxt+3): always sentinels
  auto load_slot = [&](int i) -> unsigned long long {
    if (i < hc.cs) return hc.top[i];
    if (i >= hc.tail_first && i < hc.tail_first + hc.tail_n) return hc.tail[i - hc.tail_first];
    return __ldcg(A + i);
  };
  auto store_slot = [&](int i, unsigned long long v) {
    __stcg(A + i, v);
    if (i < hc.cs) hc.top[i] = v;
    else if (i >= hc.tail_first && i < hc.tail_first + hc.tail_n) hc.tail[i - hc.tail_first] = v;
  };
  bool act = false;
  int slot = 0, cur = cap, my_x = 0;                        // hole index, pair index of its children (slots 2cur, 2cur+1)
  unsigned long long s = sent;
  int next_x = 0, wait = 0;
  unsigned ticks = 0, stalls = 0;
  if (lane == 0 && extract > 0) outs[0] = load_slot(1);
  unsigned long long nxt = load_slot(n);
  while (true) {
    // (1) children pair of the hole
    ulonglong2 pr = make_ulonglong2(sent, sent);
    if (act) {
      const int i = 2 * cur;
      if (i + 1 < hc.cs) pr = *reinterpret_cast<const ulonglong2 *>(hc.top + i);
      else if (i >= hc.tail_first && i + 1 < hc.tail_first + hc.tail_n) pr = make_ulonglong2(hc.tail[i - hc.tail_first], hc.tail[i + 1 - hc.tail_first]);
      else pr = ldcg_pair(A + i);
    }
    // (2) fill the hole, move one level down or end
    if (act) {
      const float xv = hval(pr.x), yv = hval(pr.y), sv = hval(s);
      const bool right = hcmp<MAXHEAP>(xv, yv);
      const unsigned long long c = right ? pr.y : pr.x;
      const float cv = hval(c);
      const bool stop = hstop<MAXHEAP>(sv, cv) || (MAXHEAP && cv < lose_below);
      const unsigned long long put = stop ? s : c;
      store_slot(slot, put);
      if (slot == 1) outs[my_x + 1] = put;
      if (stop) act = false;
      else { slot = 2 * cur + (right ? 1 : 0); cur = min(slot, cap); }
    }
    __syncwarp();
    // (3) may the next extraction start?  (holes as they are after the move)
    if (--wait <= 0) {
      if (next_x >= extract) { if (!__any_sync(FULL, act)) break; }
      else {
        const int ms = n - next_x;
        const int ln = next_x & (NL - 1);
        bool blocks = act && ((int)lane == ln);
        // a tail slot holding a loser is never a hole and never decides anything (loser cut), and a loser stays a loser
        if (!(MAXHEAP && hval(nxt) < lose_below)) {
          const int dh = 31 - __clz(max(slot, 1)), dms = 31 - __clz(ms);
          blocks = blocks || (act && dms >= dh && (ms >> (dms - dh)) == slot);
        }
        if (!__any_sync(FULL, blocks)) {
          if ((int)lane == ln) {
            s = load_slot(ms);
            store_slot(ms, sent);
            my_x = next_x; act = true; slot = 1; cur = 1;
          }
          next_x++; wait = 2;
          nxt = load_slot(n - next_x);                    // the next tail slot's content (every lane: a broadcast)
        } else stalls++;
      }
    }
    __syncwarp();
    ticks++;
  }
  ticks_out = ticks; stalls_out = stalls;
}

__device__ __forceinline__ void lds_pair(unsigned addr, unsigned &x0, unsigned &x1, unsigned &y0, unsigned &y1) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(y0), "=r"(y1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void lds_one(unsigned addr, unsigned &x0, unsigned &x1) {
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x0), "=r"(x1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts_one(unsigned addr, unsigned x0, unsigned x1) {
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(addr), "r"(x0), "r"(x1) : "memory");
}

// Fast single-thread extraction replay.  Same comparisons as heap_extract_seq, arranged so that the
// loop-carried dependence of one tree level is  LDS.128 -> compare -> select next address -> LDS.128 :
//   * a freed tail slot is overwritten with a sentinel (-inf for the max-heap, +inf for the min-heap) and
//     so is everything between n+1 and the last child slot (heap_pad_sentinels), which makes both bounds
//     tests of the reference loop ("child <= n", "child < n") fall out of the value comparisons: a
//     missing right child never wins, a missing pair stops the sift;
//   * the children of BOTH possible next parents have known addresses before the comparison resolves, so
//     the next pair is requested right after the compare, ahead of the stop test (a harmless read when
//     the sift ends; addresses are clamped to a sentinel pair at the end of the array).  ptxas sinks a
//     load below the branch when nothing on the exit path reads it, and ld.volatile costs ~50 cycles per
//     level (tools/ubench/heapx.cu: 97 vs 54 cycles/level), hence the `sink` value that the exit path
//     folds into the statistics word;
//   * the extracted roots go to `outv` (k-th extracted at outv[k]) instead of the tail slots.
// ~45 instead of ~100 cycles per level.  Ends with a barrier.
// work for the otherwise idle warps during a single-thread extraction replay: reset the per-node slots of the
// tokens created in this frame (clear_tokens, beam.c:1122, moved ahead of the next frame)
struct SlotClear {
  const Tok *tok; int n; SlotView slots;
  __device__ __forceinline__ void run(int first, int stride) const {
    for (int i = first; i < n; i += stride) {
      const int node = tok[i].node;
      slots.reset(node);
    }
  }
};

template <bool MAXHEAP>
__device__ __forceinline__ void heap_pad_sentinels(unsigned long long *A, const int n, const int maxt) {
  const unsigned long long sent = MAXHEAP ? 0xff800000ull : 0x7f800000ull;
  const int hi = min(2 * n + 1, maxt + 1);
  for (int i = n + 1 + (int)threadIdx.x; i <= hi; i += BEAM_THREADS) A[i] = sent;
  if (threadIdx.x < 2) A[maxt + 2 + threadIdx.x] = sent;
}

template <bool MAXHEAP>
__device__ __forceinline__ unsigned heap_pick(unsigned x0, unsigned y0, unsigned a_left, unsigned a_right, bool &right) {
  // "child < child+1" (beam.c:1359 / :1425) and the address of the chosen child's own children, one select
  unsigned r, pr;
  if (MAXHEAP) asm("{ .reg .pred p; setp.lt.f32 p, %4, %5; selp.u32 %0, %2, %3, p; selp.u32 %1, 1, 0, p; }"
                   : "=r"(r), "=r"(pr) : "r"(a_right), "r"(a_left), "f"(__uint_as_float(x0)), "f"(__uint_as_float(y0)));
  else asm("{ .reg .pred p; setp.gt.f32 p, %4, %5; selp.u32 %0, %2, %3, p; selp.u32 %1, 1, 0, p; }"
           : "=r"(r), "=r"(pr) : "r"(a_right), "r"(a_left), "f"(__uint_as_float(x0)), "f"(__uint_as_float(y0)));
  right = (pr != 0u);
  return r;
}

// one tree level: pair (X0,X1),(Y0,Y1) = children of the parent at `slot`; requests the next pair into N*
#define JB_HEAP_LEVEL(X0, X1, Y0, Y1, N0, N1, N2, N3)                                                        \
  {                                                                                                          \
    const unsigned u = (cur << 1) - hb;                                                                      \
    bool right;                                                                                              \
    const unsigned ncur = heap_pick<MAXHEAP>(X0, Y0, min(u, capa), min(u + 16u, capa), right);               \
    lds_pair(ncur, N0, N1, N2, N3);                      /* speculative: children of the chosen child */     \
    sink = N0;                                           /* (read on the exit path too, see below) */        \
    const unsigned c_lo = right ? Y0 : X0, c_hi = right ? Y1 : X1;                                           \
    levels++;                                                                                                \
    if (hstop<MAXHEAP>(sv, __uint_as_float(c_lo)) || (MAXHEAP && __uint_as_float(c_lo) < lose_below)) break; \
    sts_one(slot, c_lo, c_hi);                                                                               \
    slot = cur + (right ? 8u : 0u);                                                                          \
    cur = ncur;                                                                                              \
  }

template <bool MAXHEAP>
__device__ void heap_extract_fast(unsigned long long *A, const int n, const int extract, const float lose_below,
                                  unsigned long long *outv, const int maxt, unsigned long long *stats,
                                  const SlotClear *idle_work = nullptr, const int single_thread = 0,
                                  unsigned long long *gcache = nullptr, const int gcache_n = 0,
                                  unsigned long long *gtail = nullptr, const int gtail_n = 0) {
  HeapCache hc{nullptr, 0, nullptr, 0, 0};
  if (!__isShared(A) && gcache && n + 8 <= gcache_n && single_thread != 1) {
    // global-memory heap that fits the shared-memory area as a whole (select #1 of a multipath frame, and the smaller
    // frames of a wide beam): replay on a shared-memory copy at shared-memory speed, then write the arrangement back
    const int lmaxt = (gcache_n - 4) & ~1;
    for (int i = threadIdx.x; i <= n; i += BEAM_THREADS) gcache[i] = A[i];
    heap_pad_sentinels<MAXHEAP>(gcache, n, lmaxt);
    __syncthreads();
    if (threadIdx.x >= 32 && idle_work) idle_work->run((int)threadIdx.x - 32, BEAM_THREADS - 32);
    if (threadIdx.x < 32) {
      unsigned ticks, stalls;
      heap_extract_pipe_warp6<MAXHEAP, 0>(gcache, n, extract, lose_below, outv, lmaxt, threadIdx.x, ticks, stalls);
      if (threadIdx.x == 0) {
        atomicAdd(stats + 1, (unsigned long long)ticks); atomicAdd(stats + 2, (unsigned long long)extract);
        atomicAdd(stats + 3, (unsigned long long)stalls);
      }
    }
    __syncthreads();
    for (int i = 1 + threadIdx.x; i <= n; i += BEAM_THREADS) A[i] = gcache[i];
    __syncthreads();
    return;
  }
  if (!__isShared(A) && gcache && gcache_n > 0) {
    // global-memory heap: copy its top levels and its tail into shared memory first (all threads)
    hc.top = gcache; hc.cs = min(gcache_n, maxt + 4) & ~1;
    for (int i = threadIdx.x; i < hc.cs; i += BEAM_THREADS) gcache[i] = A[i];
    if (gtail && gtail_n >= extract) {
      hc.tail = gtail; hc.tail_first = n - extract + 1; hc.tail_n = extract;
      for (int i = threadIdx.x; i < extract; i += BEAM_THREADS) gtail[i] = A[hc.tail_first + i];
    }
    __syncthreads();
  }
  if (threadIdx.x >= 32 && idle_work) idle_work->run((int)threadIdx.x - 32, BEAM_THREADS - 32);
  if (single_thread != 1 || !__isShared(A)) {
    // warp 0: up to 16 extractions in flight, one tree level per tick each (heap_pipe.cuh)
    if (threadIdx.x < 32) {
      unsigned ticks, stalls;
      if (!__isShared(A)) heap_extract_pipe_global<MAXHEAP>(A, n, extract, lose_below, outv, maxt, hc, ticks, stalls);
      else
      if (single_thread == 2) heap_extract_pipe_warp<MAXHEAP>(A, n, extract, lose_below, outv, maxt, threadIdx.x, ticks, stalls);
      else if (single_thread == 3) heap_extract_pipe_warp4<MAXHEAP, 0>(A, n, extract, lose_below, outv, maxt, threadIdx.x, ticks, stalls);
      else heap_extract_pipe_warp6<MAXHEAP, 0>(A, n, extract, lose_below, outv, maxt, threadIdx.x, ticks, stalls);
      if (threadIdx.x == 0) {
        atomicAdd(stats + 1, (unsigned long long)ticks); atomicAdd(stats + 2, (unsigned long long)extract);
        atomicAdd(stats + 3, (unsigned long long)stalls);
      }
    }
    __syncthreads();
    return;
  }
  if (threadIdx.x == 0) {
    unsigned levels = 0, sink = 0, sinkacc = 0;
    const unsigned hb = smem_u32(A);
    const unsigned sent = MAXHEAP ? 0xff800000u : 0x7f800000u;
    const unsigned capa = hb + (((unsigned)(maxt >> 1) + 1u) << 4);    // pair (maxt+2, maxt+3): always sentinels
    unsigned mslot = hb + ((unsigned)n << 3);
    for (int x = 0; x < extract; x++) {
      unsigned s_lo, s_hi, r_lo, r_hi, x0, x1, y0, y1, z0, z1, w0, w1;
      lds_one(mslot, s_lo, s_hi);                       // s = A[m]
      sts_one(mslot, sent, 0u);                         // slot m leaves the heap (before the root's children are read)
      lds_one(hb + 8u, r_lo, r_hi);                     // root
      lds_pair(hb + 16u, x0, x1, y0, y1);               // its children
      mslot -= 8u;
      outv[x] = ((unsigned long long)r_hi << 32) | r_lo;
      const float sv = __uint_as_float(s_lo);
      unsigned slot = hb + 8u, cur = hb + 16u;          // address of the parent slot / of its children pair
      while (true) {
        JB_HEAP_LEVEL(x0, x1, y0, y1, z0, z1, w0, w1)
        JB_HEAP_LEVEL(z0, z1, w0, w1, x0, x1, y0, y1)
      }
      sts_one(slot, s_lo, s_hi);
      sinkacc += sink;
    }
    atomicAdd(stats + 1, (unsigned long long)levels); atomicAdd(stats + 2, (unsigned long long)extract);
    atomicAdd(stats + 3, (unsigned long long)sinkacc);
  }
  __syncthreads();
}
#undef JB_HEAP_LEVEL

// ---- closed form of an upward select ------------------------------------------------------------------------------
// When every sift of the extraction loop ends on a loser (an element that is never extracted) an extraction is a pure
// "pull-up": the hole at the root is filled by the larger child (the left one on a tie), and so on down.  Two elements of
// equal score then keep their relative PRE-ORDER position in the tree for ever -- the one in the right subtree of their
// lowest common ancestor could only overtake by being strictly larger than everything in the left subtree -- and the
// root is first in pre-order, hence
//        extraction order = (score descending, pre-order position in the BUILT heap ascending).
// A re-inserted winner (the tail slot an extraction takes holds an element that will itself be extracted) sinks from the
// root instead and may land ahead of elements it ties with; it cannot disturb the order of anybody else.  Such an
// element e sits in a tail slot p of the built heap (tail slots are leaves, nothing is promoted into a leaf, so a tail
// slot holds its original content or an earlier extraction's s, itself a tail content) and when p is taken, at step
// k = n-p+1, the k-1 elements extracted so far and the d = depth(p) elements on the slots above p are all ahead of e.
// So if  rank(e) < k + d  for every tail element that ties with another candidate, no re-insertion can matter and the
// closed form is exact; otherwise the caller replays the loop (heap_extract_fast).  On the 20k-word workload the test
// passes in 56 % of the frames (tools/heapstat.cpp); tools/heapsim.cpp checks form + test against the plain loop.
//
// All threads call it after heap_build<true>.  Candidates = elements >= lose_below (a lower bound of the need-th largest
// score); they are sorted with a bitonic network in `keys` (shared memory: the free part of the heap array, slots
// n+1.., or a dedicated area when the heap itself lives in global memory), keys
//   [ order-preserving score key : 32 | 0xffff - pre-order position : 16 | candidate index : 16 ]  descending,
// payload (slot << 16 | token id) in `pay`.  Returns 1 and fills ordn[0..need) (visiting order = reverse extraction order)
// or returns 0 with the heap intact (everything above slot n must be padded again before a replay).
__device__ __forceinline__ int closed_subtree_size(const int c, const int n, const int H) {
  const int dc = 31 - __clz(c);
  if (dc > H) return 0;
  const int sh = H - dc;
  const int first = c << sh, width = 1 << sh;
  return (width - 1) + max(0, min(n - first + 1, width));
}

// ---- the closed form WITH re-insertions ----------------------------------------------------------------------------
// While every extraction's s is a loser the heap evolves by pull-ups and
//   (I)  slot x holds the best remaining element of subtree(x) that no ancestor of x holds,
// "best" = (score descending, pre-order position of the element's HOME slot ascending) -- which is also the extraction
// order.  A tail leaf that still holds a candidate when it is taken breaks the pure pull-up picture: the element is
// re-inserted from the root and lands on the chain of larger children where its score says, ABOVE whatever it ties with.
// (I) survives if the element's home moves to where it lands (it is at least as good as both sub-trees below it).  So the
// sorted candidate keys ARE the heap: for every flagged tail slot m, high slots first (step k = n-m+1),
//   * who sits in leaf m: walk the alive part of the order (index >= k-1) and hand out the levels of the path root..m --
//     level j goes to the first unused element whose home lies in subtree(a_j) (nested intervals of pre-order
//     positions); m holds a candidate iff level depth(m) gets one;
//   * where that element e lands: walk the order behind the root; the first element of subtree(x) is the occupant of the
//     larger child of x (the left one on a tie: that is the order); e stays at x if its score is >= that element's
//     ("STVAL >= SVAL(child)", beam.c:1362), else the hole moves into the child whose subtree holds that element's home;
//   * e's key gets the landing slot's pre-order position and moves to its place in the order (behind this step's root).
// tools/heapdyn.cpp is the CPU model (closed_dynamic_scan), exact on recorded cuts of real decodes and on random heaps
// with as few as 8 distinct scores.  One warp; returns 0 on anything unexpected (the caller replays the loop instead).
__device__ __noinline__ int closed_relocate(unsigned long long *keys, const int nc, const int n, const int need, unsigned *flags, unsigned *multi,
                                            unsigned *pay, const int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  const int H = 31 - __clz(n);
  const int tail0 = n - need + 1, fwords = (need + 31) >> 5;
  for (int w = fwords - 1; w >= 0; w--) {
    unsigned bits = flags[w];
    while (bits) {
      const int b = 31 - __clz(bits);
      const int m = tail0 + w * 32 + b;
      const int k = n - m + 1;                              // step; this step's root is keys[k-1]
      if (m <= n && m >= 2 && k <= need) {
        const int dm = 31 - __clz(m);
        const bool is_multi = (multi[(m - tail0) >> 5] >> ((m - tail0) & 31)) & 1u;    // a re-inserted element landed in this leaf too
        // --- the occupant of leaf m
        int j = 0, a = 1, lo = 0, hi = n, occ = -1;
        bool gone = false;
        for (int base = k - 1; base < nc && occ < 0 && !gone; base += 32) {
          const int idx = base + lane;
          const unsigned long long key = (idx < nc) ? keys[idx] : 0ull;
          const int pre = 0xffff - (int)((key >> 16) & 0xffffu);
          unsigned done = 0u;
          while (true) {
            const bool in = (idx < nc) && pre >= lo && pre < hi && !((done >> lane) & 1u);
            const unsigned mask = __ballot_sync(FULL, in);
            if (!mask) break;
            const int f = __ffs(mask) - 1;
            if (j == dm) { occ = base + f; break; }
            // the leaf's own candidate, pulled up to level j: only a loser can be in the leaf now
            if (!is_multi && (int)(pay[(unsigned)keys[base + f] & 0xffffu] >> 16) == m) { gone = true; break; }
            j++;
            const int nxt = m >> (dm - j);
            const int lsz = closed_subtree_size(2 * a, n, H);
            if (nxt == 2 * a) { lo = lo + 1; hi = lo + lsz; } else { lo = lo + 1 + lsz; }
            a = nxt;
            done |= (f >= 31) ? FULL : ((2u << f) - 1u);
          }
        }
        if (occ >= k) {
          // --- e = keys[occ] is re-inserted from the root of the heap of m-1 slots
          __syncwarp();
          const unsigned long long ekey = keys[occ];
          const unsigned esc = (unsigned)(ekey >> 32);
          const int msz = m - 1;
          int x = 1; lo = 0; hi = n;
          bool stop = (2 * x > msz);
          for (int base = k; base < nc && !stop; base += 32) {
            const int idx = base + lane;
            const unsigned long long key = (idx < nc) ? keys[idx] : 0ull;
            const int pre = 0xffff - (int)((key >> 16) & 0xffffu);
            unsigned done = 0u;
            while (!stop) {
              const bool in = (idx < nc) && idx != occ && pre >= lo && pre < hi && !((done >> lane) & 1u);
              const unsigned mask = __ballot_sync(FULL, in);
              if (!mask) break;
              const int f = __ffs(mask) - 1;
              const unsigned osc = __shfl_sync(FULL, (unsigned)(key >> 32), f);
              const int opre = __shfl_sync(FULL, pre, f);
              if (opre == lo) return 0;                     // an unplaced element whose home is x: cannot happen
              if (esc >= osc) { stop = true; break; }        // e stays at x
              const int lsz = closed_subtree_size(2 * x, n, H);
              if (opre < lo + 1 + lsz) { x = 2 * x; lo = lo + 1; hi = lo + lsz; }
              else { x = 2 * x + 1; lo = lo + 1 + lsz; }
              if (2 * x > msz) { stop = true; break; }
              done |= (f >= 31) ? FULL : ((2u << f) - 1u);
            }
          }
          // --- e's home is x (pre-order position lo): new key, new place among the alive elements behind this step's root
          const unsigned long long nkey = (ekey & 0xffffffff0000ffffull) | ((unsigned long long)(0xffffu - (unsigned)lo) << 16);
          // the new place is inside e's tie group: one window around occ, unless the group is wider than that
          int cnt = 0;
          {
            const int wb = max(k, occ - 16), we = wb + 31;
            const bool lo_ok = (wb == k) || ((unsigned)(keys[wb] >> 32) > esc);
            const bool hi_ok = (we >= nc) || ((unsigned)(keys[we] >> 32) < esc);
            if (lo_ok && hi_ok) {
              const int idx = wb + lane;
              const bool gt = (idx < nc) && idx != occ && keys[idx] > nkey;
              cnt = (wb - k) + __popc(__ballot_sync(FULL, gt));
            } else {
              for (int base = k; base < nc; base += 32) {
                const int idx = base + lane;
                const bool gt = (idx < nc) && idx != occ && keys[idx] > nkey;
                const unsigned mask = __ballot_sync(FULL, gt);
                cnt += __popc(mask);
                const bool le = (idx < nc) && idx != occ && !gt;
                if (__any_sync(FULL, le)) break;             // sorted descending: nothing greater further on
              }
            }
          }
          const int ins = k + cnt;
          if (ins < occ) {
            // shift keys[ins .. occ-1] up by one, from the top end
            for (int top = occ; top > ins; top -= 32) {
              const int idx = top - lane;                     // destination index
              unsigned long long v = 0ull;
              if (idx > ins) v = keys[idx - 1];
              __syncwarp();
              if (idx > ins) keys[idx] = v;
              __syncwarp();
            }
          } else if (ins > occ) {
            // shift keys[occ+1 .. ins] down by one, from the bottom end
            for (int bot = occ; bot < ins; bot += 32) {
              const int idx = bot + lane;                     // destination index
              unsigned long long v = 0ull;
              if (idx < ins) v = keys[idx + 1];
              __syncwarp();
              if (idx < ins) keys[idx] = v;
              __syncwarp();
            }
          }
          if (lane == 0) {
            keys[ins] = nkey;
            unsigned *pp = pay + ((unsigned)nkey & 0xffffu);
            *pp = ((unsigned)x << 16) | (*pp & 0xffffu);        // home slot of the candidate
            if (x >= tail0 && x < m) { flags[(x - tail0) >> 5] |= 1u << ((x - tail0) & 31); multi[(x - tail0) >> 5] |= 1u << ((x - tail0) & 31); }
          }
          __syncwarp();
        }
      }
      // next flagged slot below b in this word (a re-insertion may have flagged one)
      bits = flags[w] & ((b == 0) ? 0u : ((1u << b) - 1u));
    }
  }
  return 1;
}

__device__ int heap_select_closed(unsigned long long *heap, const int n, const int need, const float lose_below, const int maxt,
                                  unsigned long long *keys, const int key_cap, unsigned *pay, const int pay_cap,
                                  int *ordn, int *s_scratch /* [2] shared ints */, const int p_no_reloc = 0) {
  const int tid = threadIdx.x;
  int relocated = 0;
  if (n >= 65536 || maxt >= 65536 || !(lose_below > -INFINITY)) return 0;
  // tail slots (the slots the extractions take their s from) that are the home of a candidate: one bit each, at the end of pay
  // (a second bit per slot: a re-inserted element has landed there as well)
  const int tail0 = n - need + 1, fwords = (need + 31) >> 5;
  unsigned *const flags = pay + pay_cap - fwords, *const multi = pay + pay_cap - 2 * fwords;
  const bool have_flags = (2 * fwords < pay_cap);
  if (tid == 0) { s_scratch[0] = 0; s_scratch[1] = 0; }
  if (have_flags) for (int i = tid; i < 2 * fwords; i += BEAM_THREADS) multi[i] = 0u;
  __syncthreads();
  // 1. candidates (a handle per candidate from a shared counter, one atomic per warp and pass)
  const int H = 31 - __clz(n);
  const int cap = min(min(key_cap, have_flags ? pay_cap - 2 * fwords : pay_cap), 65535);   // the flag words sit behind the payload
  for (int h0 = 1; h0 <= n; h0 += BEAM_THREADS) {
    const int h = h0 + tid;
    unsigned long long e = 0ull;
    bool is_c = false;
    if (h <= n) { e = heap[h]; is_c = (hval(e) >= lose_below); }
    const unsigned m = __ballot_sync(0xffffffffu, is_c);
    int base = 0;
    if ((tid & 31) == 0 && m) base = atomicAdd(&s_scratch[0], __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (is_c) {
      const int ci = base + __popc(m & ((1u << (tid & 31)) - 1u));
      if (ci < cap) {
        // pre-order position of slot h in the complete tree of n slots
        int pre = 0, cur = 1;
        for (int b = (31 - __clz(h)) - 1; b >= 0; b--) {
          const int bit = (h >> b) & 1;
          pre += 1 + (bit ? closed_subtree_size(cur * 2, n, H) : 0);
          cur = cur * 2 + bit;
        }
        keys[ci] = ((unsigned long long)fkey(hval(e)) << 32) | ((unsigned long long)(0xffffu - (unsigned)pre) << 16) | (unsigned)ci;
        pay[ci] = ((unsigned)h << 16) | (unsigned)(e >> 32);
      }
    }
  }
  __syncthreads();
  const int nc = s_scratch[0];
  int np = 1; while (np < nc) np <<= 1;
  if (nc > cap || np > key_cap || nc < need) return 0;      // uniform: s_scratch[0] is read after the barrier
  const bool can_relocate = have_flags;
  for (int i = nc + tid; i < np; i += BEAM_THREADS) keys[i] = 0ull;
  __syncthreads();
  // 2. bitonic sort, descending
  for (int k = 2; k <= np; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < (np >> 1); i += BEAM_THREADS) {
        const int a = ((i & ~(j - 1)) << 1) | (i & (j - 1)), b = a | j;
        const unsigned long long ka = keys[a], kb = keys[b];
        const bool desc = ((a & k) == 0);
        if (desc ? (ka < kb) : (ka > kb)) { keys[a] = kb; keys[b] = ka; }
      }
      __syncthreads();
    }
  }
  // 3. the test: which tail candidates may still be in their leaf when it is taken?  If slot p still held e at step
  //    k = n-p+1, the k-1 elements extracted so far and the d = depth(p) elements above p would all be ahead of e, so
  //    rank(e) >= k + d.  The rank of an element changes only when an element of its own score is re-inserted, so an untied
  //    e with rank < k + d is gone for sure, and a tied one if even the last place of its tie group is < k + d.  The others
  //    are flagged; if none of the flagged ties with anybody no re-insertion can matter (the plain closed form is exact).
  const unsigned theta = (unsigned)(keys[need - 1] >> 32);
  for (int i = tid; i < nc; i += BEAM_THREADS) {
    const unsigned long long ki = keys[i];
    const unsigned sk = (unsigned)(ki >> 32);
    if (sk < theta) continue;
    const int slot = (int)(pay[(unsigned)ki & 0xffffu] >> 16);
    if (slot < tail0) continue;
    const bool tied = (i > 0 && (unsigned)(keys[i - 1] >> 32) == sk) || (i + 1 < nc && (unsigned)(keys[i + 1] >> 32) == sk);
    int last = i;
    if (tied) { int g = 0; while (last + 1 < nc && (unsigned)(keys[last + 1] >> 32) == sk && g < 8) { last++; g++; } if (g == 8) last = nc; }
    const int kstep = n - slot + 1, d = 31 - __clz(slot);
    if (last + 1 < kstep + d) continue;
    if (have_flags) atomicOr(flags + ((slot - tail0) >> 5), 1u << ((slot - tail0) & 31));
    if (tied) s_scratch[1] = 1;
  }
  __syncthreads();
  if (s_scratch[1]) {
    // 3b. a tied tail element may still be in its leaf when the leaf is taken: follow the few re-insertions exactly on the
    //     implicit heap (closed_relocate); warp 0, the others wait
    if (!can_relocate || p_no_reloc) return 0;
    if (tid < 32) { const int ok = closed_relocate(keys, nc, n, need, flags, multi, pay, tid); if (tid == 0) s_scratch[1] = ok ? 2 : 1; }
    __syncthreads();
    if (s_scratch[1] != 2) return 0;
    relocated = 1;
  }
  // 4. survivors in visiting order: last extracted first
  for (int k = tid; k < need; k += BEAM_THREADS) ordn[k] = (int)(pay[(unsigned)keys[need - 1 - k] & 0xffffu] & 0xffffu);
  return 1 + relocated;
}


// bt_current_max (beam.c:876-921): the best trellis word among those stored in the frame just done (raw atoms lo..hi-1,
// end time frame-1), the most recently stored one on a tie (the reference walks its list newest first and keeps the
// first maximum), traced back to the sentence start (trace_backptr, beam.c:253-301).  One thread.
__device__ void interim_best(const BeamParams &p, const int u, const jb200_atom *araw, const int lo, const int hi, const int frame) {
  UttState *st = p.state + u;
  int best = -1; float mx = JB200_LOG_ZERO;
  for (int a = hi - 1; a >= lo; a--) if (mx < araw[a].backscore) { mx = araw[a].backscore; best = a; }
  st->interim_frame = frame - 1;
  if (best < 0) { st->interim_nwords = 0; st->interim_score = JB200_LOG_ZERO; return; }
  int *w = p.interim_words + (size_t)u * MAX_WORDS;
  int n = 0, a = best;
  w[n++] = araw[a].wid;
  while (araw[a].begintime > 0) {
    a = araw[a].last;
    if (a < 0 || n >= MAX_WORDS) break;
    w[n++] = araw[a].wid;
  }
  for (int i = 0; i < n / 2; i++) { const int x = w[i]; w[i] = w[n - 1 - i]; w[n - 1 - i] = x; }
  st->interim_nwords = n; st->interim_score = mx;
}

// phase cycle accounting (thread 0 only; negligible cost)
#define PROF_MARK(k) do { if (tid == 0) { long long _n = clock64(); s_prof[k] += _n - s_tprev; s_tprev = _n; } } while (0)

// finalize_1st_pass (bt_relocate_rw + bt_sort_rw, backtrellis.c:218-267,438-478) + find_1pass_result
// (beam.c:394-424, :253-301); shared by the normal and the multipath kernel.  All threads call it.
__device__ __forceinline__ void finalize_utt(const BeamParams &p, const int u, const int tid, const int T, jb200_atom *araw, int *newidx,
                                             const int *group0, jb200_utt_result *res, int *words,
                                             int &s_natoms, int &s_overflow, int &s_found, long long &s_outbase,
                                             long long *s_prof, long long &s_tprev) {
  // ================= finalize_1st_pass: bt_relocate_rw + bt_sort_rw (backtrellis.c:218-267,438-478) ====
  // group g = atoms with end frame g (raw atoms are grouped by creation frame already);
  // inside a group order by word id (unique per group in this build: one token per node).
  const int natoms = s_natoms;
  for (int a = tid; a < natoms; a += BEAM_THREADS) {
    const jb200_atom me = araw[a];
    const int lo = group0[me.endtime], hi = group0[me.endtime + 1];
    int rank = 0;
    for (int b = lo; b < hi; b++) rank += (araw[b].wid < me.wid) ? 1 : 0;
    newidx[a] = lo + rank;
  }
  if (tid == 0) {
    if (p.atoms_in_place) s_outbase = p.atom_off[u];        // streams: each utterance keeps its own output region
    else {
      unsigned long long base = atomicAdd(p.atom_counter, (unsigned long long)natoms);
      if ((long long)(base + natoms) > p.atoms_out_cap) { s_overflow = 1; s_outbase = -1; }
      else s_outbase = (long long)base;
    }
  }
  __syncthreads();
  const long long ob = s_outbase;
  const bool can_write = (ob >= 0);
  const int kept = can_write ? natoms : 0;
  for (int a = tid; a < kept; a += BEAM_THREADS) {
    jb200_atom me = araw[a];
    me.last = (me.last < 0) ? -1 : newidx[me.last];
    p.atoms_out[ob + newidx[a]] = me;
    // find_1pass_result (beam.c:394-424): the latest end frame holding a </s> atom
    if (me.wid == p.tail_silwid && me.backscore > JB200_LOG_ZERO) atomicMax(&s_found, me.endtime);
  }
  __threadfence_block();
  __syncthreads();

  if (tid == 0) {
    int status = 0, nw = 0; float score = 0.0f;
    const int last_time = s_found;
    if (kept == 0 || last_time < 0) status = -1;
    else {
      // the unique </s> atom of group last_time
      const int lo = group0[last_time], hi = group0[last_time + 1];
      int best = -1;
      for (int b = lo; b < hi; b++) {
        const jb200_atom x = p.atoms_out[ob + b];
        if (x.wid == p.tail_silwid && x.backscore > JB200_LOG_ZERO) { best = b; break; }
      }
      if (best < 0) status = -1;
      else {
        // trace_backptr (beam.c:253-301)
        int tmp[MAX_WORDS]; int n = 0; int a = best;
        tmp[n++] = p.atoms_out[ob + a].wid;
        while (p.atoms_out[ob + a].begintime > 0) {
          a = p.atoms_out[ob + a].last;
          if (a < 0 || n >= MAX_WORDS) break;
          tmp[n++] = p.atoms_out[ob + a].wid;
        }
        for (int i = 0; i < n; i++) words[i] = tmp[n - i - 1];
        nw = n; score = p.atoms_out[ob + best].backscore;
      }
    }
    jb200_utt_result r;
    r.status = status; r.n_frames = T; r.n_atoms = kept; r.n_words = nw; r.score = score;
    r.atom_offset = ob; r.word_offset = u * MAX_WORDS; r.overflow = s_overflow;
    *res = r;
    PROF_MARK(7);
    if (p.prof) for (int k = 0; k < 8; k++) p.prof[(size_t)u * 8 + k] = s_prof[k];
  }
}

// The same for grammar (DFA) mode: the pass-1 result is the best atom of the last frame that holds any
// (find_1pass_result, beam.c:435-458), whatever its word.
__device__ __forceinline__ void finalize_utt_grammar(const BeamParams &p, const int u, const int tid, const int T, jb200_atom *araw, int *newidx,
                                             const int *group0, jb200_utt_result *res, int *words,
                                             int &s_natoms, int &s_overflow, int &s_found, long long &s_outbase,
                                             long long *s_prof, long long &s_tprev) {
  // ================= finalize_1st_pass: bt_relocate_rw + bt_sort_rw (backtrellis.c:218-267,438-478) ====
  // group g = atoms with end frame g (raw atoms are grouped by creation frame already);
  // inside a group order by word id (unique per group in this build: one token per node).
  const int natoms = s_natoms;
  for (int a = tid; a < natoms; a += BEAM_THREADS) {
    const jb200_atom me = araw[a];
    const int lo = group0[me.endtime], hi = group0[me.endtime + 1];
    int rank = 0;
    for (int b = lo; b < hi; b++) rank += (araw[b].wid < me.wid) ? 1 : 0;
    newidx[a] = lo + rank;
  }
  if (tid == 0) {
    if (p.atoms_in_place) s_outbase = p.atom_off[u];        // streams: each utterance keeps its own output region
    else {
      unsigned long long base = atomicAdd(p.atom_counter, (unsigned long long)natoms);
      if ((long long)(base + natoms) > p.atoms_out_cap) { s_overflow = 1; s_outbase = -1; }
      else s_outbase = (long long)base;
    }
  }
  __syncthreads();
  const long long ob = s_outbase;
  const bool can_write = (ob >= 0);
  const int kept = can_write ? natoms : 0;
  for (int a = tid; a < kept; a += BEAM_THREADS) {
    jb200_atom me = araw[a];
    me.last = (me.last < 0) ? -1 : newidx[me.last];
    p.atoms_out[ob + newidx[a]] = me;
    // find_1pass_result (beam.c:394-424): the latest end frame holding a </s> atom
    if (me.backscore > JB200_LOG_ZERO) atomicMax(&s_found, me.endtime);
  }
  __threadfence_block();
  __syncthreads();

  if (tid == 0) {
    int status = 0, nw = 0; float score = 0.0f;
    const int last_time = s_found;
    if (kept == 0 || last_time < 0) status = -1;
    else {
      // the unique </s> atom of group last_time
      const int lo = group0[last_time], hi = group0[last_time + 1];
      int best = -1; float maxscore = JB200_LOG_ZERO;
      for (int b = lo; b < hi; b++) {                      // rw[last_time][] order = word id order; strict '<' keeps the first maximum
        const jb200_atom x = p.atoms_out[ob + b];
        if (maxscore < x.backscore) { maxscore = x.backscore; best = b; }
      }
      if (best < 0) status = -1;
      else {
        // trace_backptr (beam.c:253-301)
        int tmp[MAX_WORDS]; int n = 0; int a = best;
        tmp[n++] = p.atoms_out[ob + a].wid;
        while (p.atoms_out[ob + a].begintime > 0) {
          a = p.atoms_out[ob + a].last;
          if (a < 0 || n >= MAX_WORDS) break;
          tmp[n++] = p.atoms_out[ob + a].wid;
        }
        for (int i = 0; i < n; i++) words[i] = tmp[n - i - 1];
        nw = n; score = p.atoms_out[ob + best].backscore;
      }
    }
    jb200_utt_result r;
    r.status = status; r.n_frames = T; r.n_atoms = kept; r.n_words = nw; r.score = score;
    r.atom_offset = ob; r.word_offset = u * MAX_WORDS; r.overflow = s_overflow;
    *res = r;
    PROF_MARK(7);
    if (p.prof) for (int k = 0; k < 8; k++) p.prof[(size_t)u * 8 + k] = s_prof[k];
  }
}

// ---- the kernel ------------------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char beam_smem[];

#ifndef JB200_BEAM_MINBLOCKS
#define JB200_BEAM_MINBLOCKS 4
#endif
#define BEAM_KERNEL_NAME beam_kernel
#define BEAM_GRAMMAR 0
#include "beam_frames.inc"
#undef BEAM_KERNEL_NAME
#undef BEAM_GRAMMAR
#define BEAM_KERNEL_NAME beam_kernel_grammar
#define BEAM_GRAMMAR 1
#include "beam_frames.inc"
#undef BEAM_KERNEL_NAME
#undef BEAM_GRAMMAR

// ---- the multipath kernel ----------------------------------------------------------------------------
// get_back_trellis_proceed, MULTIPATH branch (beam.c:2752-2828, :2930-2941).  Trees of multipath models
// carry non-emitting word-begin / word-end nodes, and a frame runs in two halves:
//   A  word-internal transitions of the survivors of t-1 (no output probability yet), then the beam cut
//      (heap select #1) on the bare transition scores;
//   B  the word-end tokens among THOSE survivors are stored as trellis words and expanded across words
//      into the same frame's token set (onto the successors of the roots); then the output probabilities
//      of all emitting tokens are added and the beam is cut again (heap select #2).
// Select #2 runs on the token index array exactly as select #1 left it (remaining heap + extracted tail)
// with the tokens of half B appended, so select #1 is replayed in full, in place -- no loser cut there.
// Half B reuses the per-node slots: a token made in half A keeps the slot with firstseq = id - 2^30 (< 0:
// "exists") and bestkey = (score, seq 0), so later arrivals only replace its content when strictly better.
template <bool MAXHEAP>
__device__ __forceinline__ int select_exact(unsigned long long *heap, int n, int need, int *ordn, unsigned long long *outv, int maxt,
                                            unsigned long long *stats, const int heap_single,
                                            unsigned long long *gcache, const int gcache_n, unsigned long long *gtail, const int gtail_n) {
  // sort_token_no_order (beam.c:1492-1520) replayed in full; the extracted roots are put back into the
  // tail slots where the in-place algorithm leaves them (k-th extracted at slot n-k).  Returns the first
  // survivor's slot.
  const int extract = MAXHEAP ? need : n - need;
  heap_pad_sentinels<MAXHEAP>(heap, n, maxt);
  heap_build<MAXHEAP>(heap, n);
  heap_extract_fast<MAXHEAP>(heap, n, extract, -INFINITY, outv, maxt, stats, nullptr, heap_single, gcache, gcache_n, gtail, gtail_n);
  for (int k = threadIdx.x; k < extract; k += BEAM_THREADS) heap[n - k] = outv[k];
  __syncthreads();
  const int start = MAXHEAP ? n - need : 0;
  for (int k = threadIdx.x; k < need; k += BEAM_THREADS) ordn[k] = (int)(heap[start + k + 1] >> 32);
  return start;
}

static constexpr int TOK_EXISTS = 0x40000000;

__global__ void __launch_bounds__(BEAM_THREADS, JB200_BEAM_MINBLOCKS)
beam_kernel_mp(const BeamParams p) {
  const int u = blockIdx.x;
  const int tid = threadIdx.x;
  // this launch covers frames [ck.t0, ck.t1) of the utterance (ChunkDesc); f_begin only addresses its work areas
  const ChunkDesc ck = p.chunk[u];
  if (ck.flags & CHUNK_SKIP) return;
  const bool ck_first = (ck.flags & CHUNK_FIRST) != 0, ck_final = (ck.flags & CHUNK_FINAL) != 0;
  UttState *const ust = p.state + u;
  const int f_begin = p.frame_off[u];
  const int T = ck.t1;                               // frames so far; the utterance's length when ck_final
  const int MAXT = p.maxt, MAXC = p.maxc, MAXW = p.maxw;

  // shared memory: [heap (MAXT+4 entries) | offs], or, when the heap lives in global memory, [sort area | offs]
  unsigned long long *const smem_q = reinterpret_cast<unsigned long long *>(beam_smem);
  unsigned long long *heap = p.heap_g ? p.heap_g + (size_t)blockIdx.x * (MAXT + 4) : smem_q;
  int *offs = reinterpret_cast<int *>(smem_q + (p.heap_g ? p.qcap : MAXT + 4));
  int *hist = offs;                                                               // reused by select #2
  // global-memory heap: the area in front of offs doubles as the replay's copy of the heap's top levels, and a copy of
  // the tail slots follows offs
  unsigned long long *const gq = p.heap_g ? smem_q : nullptr;
  const int gqn = p.heap_g ? p.qcap : 0;
  unsigned long long *const gtail = p.heap_g ? reinterpret_cast<unsigned long long *>(offs + 2 * (p.beam + 2)) : nullptr;
  const int gtn = p.heap_g ? p.beam + 1 : 0;
  __shared__ int s_warp[NWARP + 1];
  __shared__ int s_E, s_natoms, s_ns, s_cur, s_overflow, s_found, s_cf[2];
  __shared__ unsigned s_pmaxkey, s_hmaxkey, s_losekey;
  __shared__ unsigned long long s_webest;
  __shared__ float s_thr;
  __shared__ long long s_outbase;
  __shared__ long long s_prof[8], s_tprev;

  Tok *tok0 = p.tok + (size_t)u * 2 * MAXT;
  int *ord0 = p.order + (size_t)u * 2 * MAXT;
  const SlotView slots{p.slots + (size_t)u * p.n_nodes};
  Cand *cand = p.cand + (size_t)u * MAXC;
  IsoCand *iso = p.iso + (size_t)u * max(max(p.n_iso, p.n_isoarc), 1);
  WEnd *wend = p.wend + (size_t)u * MAXW;
  unsigned *bits = p.bitmask + (size_t)u * (p.maxbits >> 5);
  int *wpre = p.wordpre + (size_t)u * (p.maxbits >> 5);
  unsigned long long *outv = reinterpret_cast<unsigned long long *>(offs);   // [beam+1] extracted roots; offs is dead during the selects
  const long long a0 = p.atom_off[u];
  const int atom_cap = (int)(p.atom_off[u + 1] - a0);
  jb200_atom *araw = p.atoms_raw + a0;
  int *newidx = p.newidx + a0;
  int *group0 = p.group0 + (size_t)f_begin + u;
  int *counts = p.counts + (size_t)f_begin * 2;
  jb200_utt_result *res = p.results + u;
  int *words = p.words + (size_t)u * MAX_WORDS;

  if (tid == 0) {
    s_found = -1; s_tprev = clock64();
    if (ck_first) { s_natoms = 0; s_overflow = 0; s_thr = JB200_LOG_ZERO; s_cur = 0; s_ns = 0; for (int k = 0; k < 8; k++) s_prof[k] = 0; }
    else { s_natoms = ust->natoms; s_overflow = ust->overflow; s_thr = ust->thr; s_cur = ust->cur; s_ns = ust->ns; for (int k = 0; k < 8; k++) s_prof[k] = ust->prof[k]; }
  }
  __syncthreads();

  // init_nodescore (beam.c:1631-1665): the word-begin node of <s> has no output (:1654-1656)
  if (T > 0 && ck_first && tid == 0) {
    const int node = p.head_node;
    const NodeRec nr = p.nodes[node];
    Tok tk;
    float ll = (nr.scid != 0) ? max_successor_prob(p, -1, nr.scid) : 0.0f;
    ll = ll * p.lm_weight + p.lm_penalty;
    tk.lscore = ll; tk.tre = -1; tk.cword = -1; tk.tre_wid = -1; tk.node = node; tk.score = ll;
    tok0[0] = tk;
    ord0[0] = 0;
    s_ns = 1;
  }
  __syncthreads();

  int tnum_prev = ck_first ? ((T > 0) ? 1 : 0) : ust->tnum_prev;
  int stopped = ck_first ? -1 : ust->stopped;       // frame at which the beam ran empty (beam.c:3012-3015), -1 = alive
  int n_left = 0;             // tokens of the unfinished (final) frame whose node slots are still set
  bool slots_clean = ck_first ? false : (ust->slots_clean != 0);   // the previous frame's node slots were already reset under its second beam cut

  // frames 0..T-1 (pass1.c:239-245 calls proceed(0) right after init), then proceed(T, final) (beam.c:3066-3072)
  const int t_last = ck_final ? T : T - 1;
  for (int t = ck.t0; t <= t_last && T > 0 && stopped < 0; t++) {
    const bool final = (t == T);
    const int cur = s_cur, nxt = cur ^ 1;
    Tok *tl = tok0 + (size_t)cur * MAXT, *tn = tok0 + (size_t)nxt * MAXT;
    int *ordl = ord0 + (size_t)cur * MAXT, *ordn = ord0 + (size_t)nxt * MAXT;
    const int ns = s_ns;
    const float thr = s_thr;
    const float *row = final ? p.rows : p.rows + (size_t)((long long)ck.row_base + t) * p.row_stride;   // the final half frame reads no scores

    // ---- P0: clear_tokens (normally done already under the previous frame's select #2)
    if (!slots_clean) {
      for (int i = tid; i < tnum_prev; i += BEAM_THREADS) {
        const int node = tl[i].node;
        slots.reset(node);
      }
    }
    slots_clean = false;
    if (tid == 0) { s_webest = 0ull; s_pmaxkey = fkey(JB200_LOG_ZERO); s_hmaxkey = 0u; if (t > 0) group0[t - 1] = s_natoms; }
    __syncthreads();
    PROF_MARK(0);

    // ---- A1: candidate counts per survivor
    int cand_total;
    {
      int carry_c = 0;
      for (int j0 = 0; j0 < ns; j0 += BEAM_THREADS) {
        const int j = j0 + tid;
        int nin = 0;
        if (j < ns) {
          const Tok tk = tl[ordl[j]];
          const NodeRec nr = p.nodes[tk.node];
          if ((tk.score > JB200_LOG_ZERO) && !(tk.score < thr))
            nin = (nr.self_a != JB200_LOG_ZERO) + (nr.next_a != JB200_LOG_ZERO) + nr.arc_n;
        }
        int tot_c;
        const int oc = block_excl_scan(nin, s_warp, &tot_c);
        if (j < ns) offs[j] = carry_c + oc;
        carry_c += tot_c;
      }
      cand_total = carry_c;
      if (tid == 0) offs[ns] = carry_c;
      if (cand_total > MAXC || cand_total > p.maxbits) { if (tid == 0) s_overflow = 1; cand_total = 0; }
    }
    int nwords = (cand_total + 31) >> 5;
    for (int w = tid; w < nwords; w += BEAM_THREADS) bits[w] = 0u;
    __syncthreads();
    PROF_MARK(1);

    // ---- A2: word-internal transitions (beam_intra_word(_core), beam.c:2004-2177), one thread per
    //           candidate (survivors near the tree roots fan out 10-20 ways: per-survivor loops leave most
    //           of the block idle); the owner of candidate c is found by bisection of the offsets
    for (int c = tid; c < cand_total; c += BEAM_THREADS) {
      int lo = 0, hi = ns;
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offs[mid] <= c) lo = mid; else hi = mid; }
      const int j = lo;
      int k = c - offs[j];
      const Tok tk = tl[ordl[j]];
      const NodeRec nr = p.nodes[tk.node];
      int next; float pa;
      const int has_self = (nr.self_a != JB200_LOG_ZERO), has_next = (nr.next_a != JB200_LOG_ZERO);
      if (has_self && k == 0) { next = tk.node; pa = nr.self_a; }
      else if (has_next && k == has_self) { next = nr.next; pa = nr.next_a; }
      else { const int a = k - has_self - has_next; next = __ldg(p.arc_to + nr.arc_off + a); pa = __ldg(p.arc_a + nr.arc_off + a); }
      float tmpsum = tk.score + pa;
      float lsc = JB200_LOG_ZERO;
      if (next != tk.node) {
        const int scid = p.nodes[next].scid;
        if (scid != 0) {
          lsc = max_successor_prob(p, tk.cword, scid) * p.lm_weight + p.lm_penalty;
          tmpsum -= tk.lscore;
          tmpsum += lsc;
        }
      }
      if (lsc == JB200_LOG_ZERO) lsc = tk.lscore;
      Cand cd; cd.score = tmpsum; cd.node = next; cd.lscore = lsc; cd.src = j;
      cand[c] = cd;
      if (tmpsum > JB200_LOG_ZERO) {
        const unsigned seq = (unsigned)j * SEQ_LOCAL + (unsigned)k;
        cand_atomics(slots, next, tmpsum, seq, seq);
      }
    }
    __syncthreads();
    PROF_MARK(2);

    // ---- A3: creators, in arrival order = candidate order
    for (int c = tid; c < cand_total; c += BEAM_THREADS) {
      const Cand cd = cand[c];
      if (!(cd.score > JB200_LOG_ZERO)) continue;
      const unsigned seq = (unsigned)cd.src * SEQ_LOCAL + (unsigned)(c - offs[cd.src]);
      if ((unsigned)__ldcg(slots.fs(cd.node)) == seq) atomicOr(bits + (c >> 5), 1u << (c & 31));
    }
    __syncthreads();
    PROF_MARK(3);
    int ncre_a;
    {
      int carry = 0;
      for (int w0 = 0; w0 < nwords; w0 += BEAM_THREADS) {
        const int w = w0 + tid;
        const int cnt = (w < nwords) ? __popc(__ldcg(bits + w)) : 0;
        int tot;
        const int ex = block_excl_scan(cnt, s_warp, &tot);
        if (w < nwords) wpre[w] = carry + ex;
        carry += tot;
      }
      ncre_a = carry;
    }
    if (ncre_a > MAXT) { if (tid == 0) s_overflow = 1; ncre_a = 0; }
    __syncthreads();
    PROF_MARK(4);

    // ---- A4: materialise the tokens of half A (no output probability yet) and mark their slots "exists"
    for (int c = tid; c < cand_total && ncre_a > 0; c += BEAM_THREADS) {
      const unsigned wbits = __ldcg(bits + (c >> 5));
      if (!((wbits >> (c & 31)) & 1u)) continue;
      const int r = wpre[c >> 5] + __popc(wbits & ((1u << (c & 31)) - 1u));
      const int node = cand[c].node;
      const unsigned long long bk = __ldcg(slots.bk(node));
      const unsigned seqw = ~(unsigned)(bk & 0xffffffffu);
      const int j = (int)(seqw >> SEQ_LOCAL_BITS), local = (int)(seqw & (SEQ_LOCAL - 1));
      const Cand cd = cand[offs[j] + local];
      const Tok src = tl[ordl[j]];
      Tok nt; nt.node = node;
      nt.score = cd.score; nt.lscore = cd.lscore; nt.tre = src.tre; nt.cword = src.cword; nt.tre_wid = src.tre_wid;
      tn[r] = nt;
      heap[r + 1] = ((unsigned long long)(unsigned)r << 32) | __float_as_uint(nt.score);
      slots.set(node, r - TOK_EXISTS, ((unsigned long long)fkey(nt.score) << 32) | 0xffffffffull);
    }
    __syncthreads();
    PROF_MARK(5);

    // ---- A5: heap select #1, replayed in full
    int ns_a;
    {
      const int need = p.beam;
      if (need >= ncre_a) {
        ns_a = ncre_a;
        for (int k = tid; k < ns_a; k += BEAM_THREADS) ordn[k] = k;
      } else {
        ns_a = need;
        if (need < ncre_a - need) select_exact<true>(heap, ncre_a, need, ordn, outv, MAXT, p.misspec_counter, p.heap_single, gq, gqn, gtail, gtn);
        else select_exact<false>(heap, ncre_a, need, ordn, outv, MAXT, p.misspec_counter, p.heap_single, gq, gqn, gtail, gtn);
      }
    }
    __syncthreads();
    PROF_MARK(6);

    // ---- B1: word ends among the survivors of select #1: trellis words (save_trellis, beam.c:2209)
    //          and the cross-word sources (beam_inter_word's per-token part, beam.c:2271-2335)
    int nbits_b;
    {
      int carry_a = s_natoms, carry_w = 0;
      for (int k0 = 0; k0 < ns_a; k0 += BEAM_THREADS) {
        const int k = k0 + tid;
        int is_we = 0, is_tr = 0;
        Tok tk; NodeRec nr;
        if (k < ns_a) {
          tk = tn[ordn[k]];
          nr = p.nodes[tk.node];
          if (!(tk.score < thr) && nr.stend >= 0) { is_we = 1; is_tr = (!final && nr.stend != p.tail_silwid); }
        }
        int tot_a, tot_w;
        const int oa = block_excl_scan(is_we, s_warp, &tot_a);
        const int ow = block_excl_scan(is_tr, s_warp, &tot_w);
        if (is_we) {
          const int ai = carry_a + oa;
          if (ai < atom_cap && t > 0) {
            jb200_atom a;
            a.wid = nr.stend; a.backscore = tk.score;
            a.begintime = (tk.tre < 0 ? -1 : araw[tk.tre].endtime) + 1;
            a.endtime = t - 1; a.last = tk.tre; a.lscore = tk.lscore;
            araw[ai] = a;
          } else s_overflow = 1;
          if (is_tr) {
            const int wi = carry_w + ow;
            if (wi < MAXW && ai < atom_cap) {
              WEnd w;
              const int sword = nr.stend;
              const int transp = p.is_transp[sword];
              w.j = k; w.atom = ai; w.last_word = transp ? tk.cword : sword;
              w.base = tk.score;                                   // no wordend_a in multipath (beam.c:2307)
              w.transp2 = (transp && tk.cword >= 0 && p.is_transp[tk.cword]) ? 1 : 0;
              w.nintra = 0;
              wend[wi] = w;
              if (w.base > JB200_LOG_ZERO)
                atomicMax(&s_webest, ((unsigned long long)fkey(w.base) << 32) | (unsigned)(~(unsigned)wi));
            } else s_overflow = 1;
          }
        }
        carry_a += tot_a; carry_w += tot_w;
      }
      if (tid == 0) { s_natoms = min(carry_a, atom_cap); s_E = min(carry_w, MAXW); }
      nbits_b = carry_w * p.n_isoarc + p.n_sharc;
      if (carry_w > MAXW || nbits_b > p.maxbits) { if (tid == 0) s_overflow = 1; nbits_b = 0; }
    }
    if (final) { n_left = ncre_a; __syncthreads(); break; }
    nwords = (nbits_b + 31) >> 5;
    for (int w = tid; w < nwords; w += BEAM_THREADS) bits[w] = 0u;
    __syncthreads();
    PROF_MARK(1);
    const int E = (nbits_b > 0) ? s_E : 0;

    // ---- B2: cross-word transitions through the isolated roots (beam.c:2336-2500), one candidate per
    //          (word end, root successor), pre-reduced per successor over the word ends in visiting order
    for (int ia = tid; ia < p.n_isoarc; ia += BEAM_THREADS) {
      const int col = __ldg(p.iso_id + __ldg(p.isoarc_iso + ia));
      const float pa = __ldg(p.isoarc_a + ia);
      float best = JB200_LOG_ZERO, bestl = 0.0f; int beste = -1, firste = -1;
      for (int e = 0; e < E; e++) {
        const WEnd w = wend[e];
        const float tmpprob = __ldg(p.iw + (size_t)w.last_word * p.n_iso + col);
        const float lsc = tmpprob * p.lm_weight + p.lm_penalty;
        float tmpsum = w.base;
        tmpsum += lsc;
        if (w.transp2) tmpsum += p.lm_penalty_trans;
        const float v = tmpsum + pa;
        if (v > JB200_LOG_ZERO) {
          if (firste < 0) firste = e;
          if (beste < 0 || best < v) { best = v; beste = e; bestl = lsc; }
        }
      }
      IsoCand ic; ic.score = best; ic.e = beste; ic.lscore = bestl; ic.first_e = firste;
      iso[ia] = ic;
      if (firste >= 0) {
        const unsigned sf = (unsigned)(wend[firste].j + 1) * SEQ_LOCAL + (unsigned)ia;
        const unsigned sw = (unsigned)(wend[beste].j + 1) * SEQ_LOCAL + (unsigned)ia;
        cand_atomics(slots, __ldg(p.isoarc_node + ia), best, sf, sw);
      }
    }
    // ---- B3: best word end -> successors of the shared (1-gram factored) roots (beam.c:2549-2616)
    const unsigned long long webest = s_webest;
    const bool have_we = (webest != 0ull) && (nbits_b > 0);
    WEnd wbest; wbest.base = 0.0f; wbest.atom = -1; wbest.last_word = -1; wbest.transp2 = 0; wbest.j = 0; wbest.nintra = 0;
    auto shared_value = [&](int sa, float &lsc, float &v) -> bool {
      lsc = __ldg(p.shared_f + __ldg(p.sharc_shared + sa)) * p.lm_weight + p.lm_penalty;
      float tmpsum = wbest.base;
      tmpsum += lsc;
      if (wbest.transp2) tmpsum += p.lm_penalty_trans;
      if (tmpsum < thr) return false;
      v = tmpsum + __ldg(p.sharc_a + sa);
      return v > JB200_LOG_ZERO;
    };
    if (have_we) {
      wbest = wend[(unsigned)(~(unsigned)(webest & 0xffffffffu))];
      for (int sa = tid; sa < p.n_sharc; sa += BEAM_THREADS) {
        float lsc, v;
        if (!shared_value(sa, lsc, v)) continue;
        const unsigned seq = (unsigned)(ns_a + 1) * SEQ_LOCAL + (unsigned)sa;
        cand_atomics(slots, __ldg(p.sharc_node + sa), v, seq, seq);
      }
    }
    __syncthreads();
    PROF_MARK(2);

    // ---- B4: creators of half B (arrival order: word end major, then the factoring pass)
    for (int ia = tid; ia < p.n_isoarc; ia += BEAM_THREADS) {
      const IsoCand ic = iso[ia];
      if (ic.first_e < 0) continue;
      const unsigned sf = (unsigned)(wend[ic.first_e].j + 1) * SEQ_LOCAL + (unsigned)ia;
      if ((unsigned)__ldcg(slots.fs(__ldg(p.isoarc_node + ia))) == sf) {
        const int pos = ic.first_e * p.n_isoarc + ia;
        atomicOr(bits + (pos >> 5), 1u << (pos & 31));
      }
    }
    if (have_we) {
      for (int sa = tid; sa < p.n_sharc; sa += BEAM_THREADS) {
        const unsigned seq = (unsigned)(ns_a + 1) * SEQ_LOCAL + (unsigned)sa;
        if ((unsigned)__ldcg(slots.fs(__ldg(p.sharc_node + sa))) == seq) {
          const int pos = E * p.n_isoarc + sa;
          atomicOr(bits + (pos >> 5), 1u << (pos & 31));
        }
      }
    }
    __syncthreads();
    PROF_MARK(3);
    int ncre;
    {
      int carry = 0;
      for (int w0 = 0; w0 < nwords; w0 += BEAM_THREADS) {
        const int w = w0 + tid;
        const int cnt = (w < nwords) ? __popc(__ldcg(bits + w)) : 0;
        int tot;
        const int ex = block_excl_scan(cnt, s_warp, &tot);
        if (w < nwords) wpre[w] = carry + ex;
        carry += tot;
      }
      ncre = ncre_a + carry;
    }
    if (ncre > MAXT) { if (tid == 0) s_overflow = 1; ncre = ncre_a; nbits_b = 0; }
    __syncthreads();
    PROF_MARK(4);

    // ---- B5: new tokens get the winner's content; tokens of half A that lost to a cross-word arrival
    //          are overwritten in place (propagate_token, beam.c:1901-1972)
    auto winner_content = [&](unsigned seqw, Tok &nt) {
      const int j = (int)(seqw >> SEQ_LOCAL_BITS), local = (int)(seqw & (SEQ_LOCAL - 1));
      if (j == ns_a + 1) {
        float lsc, v;
        shared_value(local, lsc, v);
        nt.score = v; nt.lscore = lsc; nt.tre = wbest.atom; nt.cword = wbest.last_word; nt.tre_wid = araw[wbest.atom].wid;
      } else {
        const IsoCand ic = iso[local];
        const WEnd w = wend[ic.e];
        nt.score = ic.score; nt.lscore = ic.lscore; nt.tre = w.atom; nt.cword = w.last_word; nt.tre_wid = araw[w.atom].wid;
      }
    };
    auto settle = [&](int node, unsigned seq_first, unsigned seq_win, int pos) {
      const int fs = __ldcg(slots.fs(node));
      const unsigned seqw = ~(unsigned)(__ldcg(slots.bk(node)) & 0xffffffffu);
      if (fs < 0) {
        if (seqw != seq_win) return;
        Tok nt; nt.node = node;
        winner_content(seqw, nt);
        tn[fs + TOK_EXISTS] = nt;
      } else if ((unsigned)fs == seq_first) {
        const unsigned wbits = __ldcg(bits + (pos >> 5));
        const int r = ncre_a + wpre[pos >> 5] + __popc(wbits & ((1u << (pos & 31)) - 1u));
        Tok nt; nt.node = node;
        winner_content(seqw, nt);
        tn[r] = nt;
      }
    };
    if (nbits_b > 0) {
      for (int ia = tid; ia < p.n_isoarc; ia += BEAM_THREADS) {
        const IsoCand ic = iso[ia];
        if (ic.first_e < 0) continue;
        settle(__ldg(p.isoarc_node + ia), (unsigned)(wend[ic.first_e].j + 1) * SEQ_LOCAL + (unsigned)ia,
               (unsigned)(wend[ic.e].j + 1) * SEQ_LOCAL + (unsigned)ia, ic.first_e * p.n_isoarc + ia);
      }
      if (have_we)
        for (int sa = tid; sa < p.n_sharc; sa += BEAM_THREADS) {
          float lsc, v;
          if (!shared_value(sa, lsc, v)) continue;
          const unsigned seq = (unsigned)(ns_a + 1) * SEQ_LOCAL + (unsigned)sa;
          settle(__ldg(p.sharc_node + sa), seq, seq, E * p.n_isoarc + sa);
        }
    }
    __syncthreads();

    // ---- B6: output probabilities of the emitting tokens (beam.c:2930-2941); then the index array select #2
    //          starts from: select #1's arrangement with fresh scores, followed by the tokens of half B
    for (int r = tid; r < ncre; r += BEAM_THREADS) {
      const Tok tk = tn[r];
      const int out = p.nodes[tk.node].out;
      float sc = tk.score;
      if (((unsigned)out >> 28) != 0xFu) {
        sc += outprob_style(p, row, out, tk.tre_wid);
        tn[r].score = sc;
        atomicMax(&s_pmaxkey, fkey(sc));
      }
      atomicMax(&s_hmaxkey, fkey(sc));
      if (r >= ncre_a) heap[r + 1] = ((unsigned long long)(unsigned)r << 32) | __float_as_uint(sc);
    }
    __syncthreads();
    for (int h = 1 + tid; h <= ncre_a; h += BEAM_THREADS) {
      const unsigned id = (unsigned)(heap[h] >> 32);
      heap[h] = ((unsigned long long)id << 32) | __float_as_uint(tn[id].score);
    }
    __syncthreads();
    PROF_MARK(5);

    // ---- B7: heap select #2 (only its survivors' order is observable: loser cut allowed)
    int ns_new;
    {
      const int need = p.beam, rest = ncre - need;
      if (need >= ncre) {
        ns_new = ncre;
        // tindex order = select #1's arrangement, then the new tokens
        for (int k = tid; k < ns_new; k += BEAM_THREADS) ordn[k] = (int)(heap[k + 1] >> 32);
      } else if (need < rest) {
        ns_new = need;
        constexpr int NB = 1024;
        const int nb = min(NB, 2 * (p.beam + 2));
        for (int i = tid; i < nb; i += BEAM_THREADS) hist[i] = 0;
        __syncthreads();
        const unsigned maxkey = s_hmaxkey;
        const int e = (int)((((maxkey & 0x80000000u) ? (maxkey & 0x7fffffffu) : ~maxkey) >> 23) & 0xffu) - 127;
        const int sh = max(0, min(24, 22 - e));
        for (int r = tid; r < ncre; r += BEAM_THREADS) {
          const unsigned key = fkey(hval(heap[r + 1]));
          const unsigned bin = min((unsigned)(nb - 1), (maxkey - key) >> sh);
          atomicAdd(&hist[bin], 1);
        }
        __syncthreads();
        if (tid < 32) {
          int cum = 0, found = -1;
          for (int b0 = 0; b0 < nb && found < 0; b0 += 32) {
            const int v = (b0 + tid < nb) ? hist[b0 + tid] : 0;
            int x = v;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (tid >= o) x += y; }
            const unsigned hit = __ballot_sync(0xffffffffu, cum + x >= need);
            if (hit) found = b0 + __ffs(hit) - 1;
            cum += __shfl_sync(0xffffffffu, x, 31);
          }
          if (tid == 0) {
            unsigned lk = 0u;
            if (found >= 0 && found < nb - 1) {
              const unsigned long long drop = (unsigned long long)(found + 1) << sh;
              lk = (drop < maxkey) ? maxkey - (unsigned)drop : 0u;
            }
            s_losekey = lk;
          }
        }
        __syncthreads();
        const unsigned lk = s_losekey;
        const float lose_below = (lk == 0u || p.no_lose) ? -INFINITY : __uint_as_float((lk & 0x80000000u) ? (lk & 0x7fffffffu) : ~lk);
        heap_build<true>(heap, ncre); PROF_MARK(7);
        const SlotClear sc{tn, ncre, slots};
        slots_clean = true;
        int closed = 0;
        if (!p.no_closed) {
          sc.run((int)threadIdx.x, BEAM_THREADS);
          closed = heap_select_closed(heap, ncre, need, lose_below, MAXT, p.heap_g ? smem_q : heap + ncre + 1, p.heap_g ? p.sort_cap : MAXT + 3 - ncre,
                                      reinterpret_cast<unsigned *>(offs), 2 * (p.beam + 2), ordn, s_cf, p.no_reloc);
          if (tid == 0) { atomicAdd(p.misspec_counter + 4, 1ull); if (closed) atomicAdd(p.misspec_counter + 5, 1ull); if (closed == 2) atomicAdd(p.misspec_counter + 6, 1ull); }
        }
        if (!closed) {
          heap_pad_sentinels<true>(heap, ncre, MAXT);
          __syncthreads();
          heap_extract_fast<true>(heap, ncre, need, lose_below, outv, MAXT, p.misspec_counter, p.no_closed ? &sc : nullptr, p.heap_single, gq, gqn, gtail, gtn);
          for (int k = tid; k < need; k += BEAM_THREADS) ordn[k] = (int)(outv[need - 1 - k] >> 32);
        }
      } else {
        ns_new = need;
        heap_pad_sentinels<false>(heap, ncre, MAXT);
        heap_build<false>(heap, ncre); PROF_MARK(7);
        const SlotClear sc{tn, ncre, slots};
        slots_clean = true;
        heap_extract_fast<false>(heap, ncre, rest, -INFINITY, outv, MAXT, p.misspec_counter, &sc, p.heap_single, gq, gqn, gtail, gtn);
        for (int k = tid; k < need; k += BEAM_THREADS) ordn[k] = (int)(heap[k + 1] >> 32);
      }
    }
    PROF_MARK(6);
    if (tid == 0) {
      counts[2 * t] = ncre; counts[2 * t + 1] = ns_new;
      s_ns = ns_new; s_cur = nxt;
      if (p.prune_width >= 0.0f) {
        const unsigned k = s_pmaxkey;
        const unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
        s_thr = __uint_as_float(b) - p.prune_width;
      } else s_thr = JB200_LOG_ZERO;
    }
    tnum_prev = ncre;
    __syncthreads();
    if (ncre == 0) { stopped = t; break; }      // beam.c:3012-3015
  }

  if (!ck_final) {
    // more frames to come: park the scalar state (everything else already lives in the utterance's global work area)
    if (tid == 0) {
      ust->natoms = s_natoms; ust->overflow = s_overflow; ust->thr = s_thr; ust->ns = s_ns; ust->cur = s_cur;
      ust->tnum_prev = tnum_prev; ust->stopped = stopped; ust->slots_clean = slots_clean ? 1 : 0; ust->n_left = 0;
      ust->t_done = T;
      long long _n = clock64(); s_prof[7] += _n - s_tprev;
      for (int k = 0; k < 8; k++) ust->prof[k] = s_prof[k];
      // the word ends of frame T-1 were stored in half B of that frame (end time T-2)
      if (p.interim) interim_best(p, u, araw, (T >= 2 && stopped < 0) ? group0[T - 2] : s_natoms, s_natoms, T - 1);
    }
    return;
  }
  const int groups = (stopped >= 0) ? stopped : T;

  {
    if (tid == 0 && T > 0) group0[groups] = s_natoms;
    // leave the node slots clean for the next utterance that uses this work area
    // (only the unfinished final frame leaves any: every other frame's slots are reset by the next P0)
    const Tok *tlast = tok0 + (size_t)(s_cur ^ 1) * MAXT;
    for (int i = tid; i < n_left; i += BEAM_THREADS) {
      const int node = tlast[i].node;
      slots.reset(node);
    }
    __syncthreads();
  }
  finalize_utt(p, u, tid, T, araw, newidx, group0, res, words, s_natoms, s_overflow, s_found, s_outbase, s_prof, s_tprev);
}

// ---- set-up kernels ------------------------------------------------------------------------------
__global__ void iw_table_kernel(BeamParams p, const int *iso_word, float *iw, int n_words) {
  // max_successor_prob_iw (factoring_sub.c:1049-1143) for EVERY last word, once
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_words * p.n_iso) return;
  const int lw = (int)(idx / p.n_iso), i = (int)(idx % p.n_iso);
  const int w = iso_word[i];
  iw[(size_t)lw * p.n_iso + p.iso_id[i]] = bigram_prob(p, p.wton[lw], p.wton[w]) + p.cprob[w];
}

__global__ void fill_slots_kernel(NodeSlot *slots, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) slots[i] = NodeSlot{0ull, 0x7fffffff, 0};
}

}  // namespace jb200

// =============================================================================================
using namespace jb200;

struct jb200_decoder {
  jb200_gmm *am = nullptr;
  jb200_dnn *dnn = nullptr;
  int device = 0, dim = 0, S = 0, row_stride = 0;
  int max_utts = 0, max_frames = 0;         // per batch: utterances, total frames
  int atoms_per_frame = 64;
  BeamParams P{};
  std::vector<void *> dev_allocs;
  // read-only tables shared by all utterances (tree, LM, inter-word table, bigram memo) sit in ONE allocation (one
  // contiguous range for an optional L2 access-policy window, see jb200_decoder_create)
  char *arena = nullptr; size_t arena_size = 0, arena_used = 0; bool l2_window = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[5]{};
  // batch buffers
  float *d_feats = nullptr, *d_rows = nullptr;
  int *d_frame_off = nullptr; long long *d_atom_off = nullptr;
  jb200_atom *d_atoms_out = nullptr; unsigned long long *d_atom_counter = nullptr;
  jb200_utt_result *d_results = nullptr; int *d_words = nullptr; long long *d_prof = nullptr;
  // host results (pinned)
  jb200_utt_result *h_results = nullptr; jb200_atom *h_atoms = nullptr; int *h_words = nullptr;
  unsigned long long *h_counter = nullptr;
  long long atoms_cap = 0;
  int last_n = 0; long long last_atoms = 0; int last_total_frames = 0;
  std::vector<int> h_frame_off;
  float last_ms[4] = {0, 0, 0, 0};
  size_t smem_bytes = 0;
  bool fetched = false;
  long long last_d2h = 0;
  int resident = 0;
  bool grammar = false;
  // chunked launches: descriptors (a pinned staging copy and its device copy, one row of max_utts per chunk), parked state
  static constexpr int MAX_CHUNKS = 64;
  ChunkDesc *h_chunk = nullptr, *d_chunk = nullptr;
  UttState *d_state = nullptr; int *d_interim_words = nullptr;
  // batch pipeline: scoring of time slice c+1 on its own stream beside the token passing of slice c
  cudaStream_t score_stream = nullptr;
  cudaEvent_t ev_slice[MAX_CHUNKS]{}; cudaEvent_t ev_score_begin = nullptr, ev_score_end = nullptr;
  int *h_seg = nullptr, *d_seg = nullptr;       // per slice: seg_off [max_utts+1] then seg_start [max_utts]
  std::vector<int> slice_nseg, slice_frames;
  int pipe_frames = 0;                          // frames per time slice; 0 = the pipeline is off
  int n_chunks = 1; float last_score_busy_ms = 0.0f; bool last_piped = false;
  // streams (jb200_stream_*)
  bool stream_mode = false; int st_n = 0, st_cap = 0;
  std::vector<int> st_t; std::vector<char> st_started, st_done;
  long long *h_aoff = nullptr;                  // pinned copy of the per-utterance atom offsets
  UttState *h_state = nullptr; int *h_interim_words = nullptr;
};

// from the shared-table arena when it has room, else an allocation of its own
static void *arena_take(jb200_decoder *d, size_t bytes) {
  const size_t need = (bytes + 255) & ~(size_t)255;
  if (!d->arena || d->arena_used + need > d->arena_size) return nullptr;
  void *p = d->arena + d->arena_used;
  d->arena_used += need;
  return p;
}
template <typename Tp>
static int dev_upload(jb200_decoder *d, const Tp *src, size_t n, const Tp **dst) {
  Tp *p = static_cast<Tp *>(arena_take(d, std::max<size_t>(n, 1) * sizeof(Tp)));
  if (!p) {
    JB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(Tp)));
    d->dev_allocs.push_back(p);
  }
  if (n) JB_CUDA(cudaMemcpy(p, src, n * sizeof(Tp), cudaMemcpyHostToDevice));
  *dst = p;
  return JB200_OK;
}
template <typename Tp>
static int dev_alloc_shared(jb200_decoder *d, size_t n, Tp **dst) {
  Tp *p = static_cast<Tp *>(arena_take(d, std::max<size_t>(n, 1) * sizeof(Tp)));
  if (!p) {
    JB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(Tp)));
    d->dev_allocs.push_back(p);
  }
  *dst = p;
  return JB200_OK;
}
template <typename Tp>
static int dev_alloc(jb200_decoder *d, size_t n, Tp **dst) {
  Tp *p = nullptr;
  JB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(Tp)));
  d->dev_allocs.push_back(p);
  *dst = p;
  return JB200_OK;
}

extern "C" void jb200_decoder_destroy(jb200_decoder *d) {
  if (!d) return;
  cudaSetDevice(d->device);
  if (d->l2_window) cudaCtxResetPersistingL2Cache();
  for (void *p : d->dev_allocs) cudaFree(p);
  if (d->h_results) cudaFreeHost(d->h_results);
  if (d->h_atoms) cudaFreeHost(d->h_atoms);
  if (d->h_words) cudaFreeHost(d->h_words);
  if (d->h_counter) cudaFreeHost(d->h_counter);
  if (d->h_chunk) cudaFreeHost(d->h_chunk);
  if (d->h_seg) cudaFreeHost(d->h_seg);
  if (d->h_aoff) cudaFreeHost(d->h_aoff);
  if (d->h_state) cudaFreeHost(d->h_state);
  if (d->h_interim_words) cudaFreeHost(d->h_interim_words);
  for (auto &e : d->ev) if (e) cudaEventDestroy(e);
  for (auto &e : d->ev_slice) if (e) cudaEventDestroy(e);
  if (d->ev_score_begin) cudaEventDestroy(d->ev_score_begin);
  if (d->ev_score_end) cudaEventDestroy(d->ev_score_end);
  if (d->score_stream) cudaStreamDestroy(d->score_stream);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
}

extern "C" int jb200_decoder_create(const jb200_tree_desc *t, jb200_gmm *am, int max_utts, int max_frames, jb200_decoder **out) {
  if (!t || !am || !out || max_utts < 1 || max_frames < 1) { set_error("jb200_decoder_create: bad argument"); return JB200_ERR_ARG; }
  const bool grammar = (t->lm_type == JB200_LM_DFA);
  if (t->lm_type != JB200_LM_NGRAM && !grammar) { set_error("unknown language-model type %d", t->lm_type); return JB200_ERR_UNSUPPORTED; }
  if (grammar) {
    if (t->multipath) { set_error("grammar mode on a multipath tree is not supported by the GPU beam"); return JB200_ERR_UNSUPPORTED; }
    if (t->n_shared != 0 || t->n_init < 1 || t->n_init > t->beam_width || !t->cp_allowed || !t->init_node || !t->init_lscore) {
      set_error("grammar mode: inconsistent descriptor (n_shared %d, n_init %d, beam %d)", t->n_shared, t->n_init, t->beam_width);
      return JB200_ERR_ARG;
    }
    for (int i = 0; i < t->n_words; i++)
      if (t->is_transparent[i]) { set_error("grammar mode: transparent words are not supported"); return JB200_ERR_UNSUPPORTED; }
  }
  if (t->n_nodes >= (1 << 28)) { set_error("lexicon tree too large"); return JB200_ERR_UNSUPPORTED; }
  // a transparent head silence would end the first word with no context word at all (last_word = -1): the reference
  // indexes wton[] / the inter-word cache with WORD_INVALID there (factoring_sub.c:1052-1056), i.e. has no defined result
  if (!grammar && (t->head_silwid < 0 || t->head_silwid >= t->n_words || t->is_transparent[t->head_silwid])) {
    set_error("the head silence word must exist and must not be transparent"); return JB200_ERR_UNSUPPORTED;
  }
  if (t->beam_width < 1 || t->beam_width > 8000) { set_error("beam width %d outside 1..8000", t->beam_width); return JB200_ERR_UNSUPPORTED; }
  jb200_decoder *d = new jb200_decoder();
  d->am = am; d->device = gmm_device(am); d->dim = gmm_dim(am);
  d->S = jb200_gmm_n_states(am);
  d->row_stride = (d->S + 3) & ~3;
  d->max_utts = max_utts; d->max_frames = max_frames;
  if (const char *e = getenv("JB200_ATOMS_PER_FRAME")) d->atoms_per_frame = std::max(4, atoi(e));
  int rc;
#define TRY(x) do { rc = (x); if (rc) { jb200_decoder_destroy(d); return rc; } } while (0)
#define TRYC(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { set_error("%s: %s", #x, cudaGetErrorString(_e)); jb200_decoder_destroy(d); return JB200_ERR_CUDA; } } while (0)
  TRYC(cudaSetDevice(d->device));
  TRYC(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  for (auto &e : d->ev) TRYC(cudaEventCreate(&e));

  BeamParams &P = d->P;
  const int n = t->n_nodes;
  {
    // shared read-only tables: nodes 32 B, arcs 8 B, context table, inter-word table, bigram memo, LM arrays (+ slack)
    const int lmc_bits_est = (t->n_words >= 65535 || t->n_scword >= 65535) ? 0 : 21;
    size_t est = (size_t)n * 32 + (size_t)t->n_arcs * 8 * 3 + (size_t)t->n_rset * (t->n_ctx + 1) * 4 + (size_t)t->n_words * 32 +
                 (size_t)t->n_words * std::max(t->n_iso, 1) * (grammar ? 1 : 4) + ((size_t)8 << lmc_bits_est) +
                 (size_t)t->lm_nvocab * 16 + (size_t)t->lm_nbigram * 8 + (size_t)(t->n_iso + t->n_shared + t->n_fscore + t->n_scword) * 16 + (4u << 20);
    if (getenv("JB200_NO_ARENA") == nullptr && cudaMalloc(&d->arena, est) == cudaSuccess) { d->arena_size = est; d->dev_allocs.push_back(d->arena); }
    else { d->arena = nullptr; cudaGetLastError(); }
  }
  // Node numbering.  The host numbers the nodes word by word (a word's own nodes are consecutive), so the ~2400 nodes a
  // frame touches are spread over the whole tree although 85 % of them sit in its first three levels (half of a frame's
  // tokens are the roots that a word end fans out to): one 128-byte line of per-node arrival slots per token.  Node ids
  // are not observable outside the decoder, so the tree is renumbered breadth-first from the roots (roots in their list
  // order, then level by level): the slots and node records a frame touches become a few dense ranges -- 3.4x fewer
  // slot lines, 2.1x fewer node-record lines per frame on the 20k-word tree (tools/node_locality.py).  `next_a` no longer
  // leads to id+1, so the record carries the successor explicitly.  JB200_NO_RENUMBER=1 keeps the host's numbering.
  std::vector<int> perm(n), inv(n);
  {
    std::vector<int> order; order.reserve(n);
    std::vector<char> seen(n, 0);
    auto push = [&](int x) { if (x >= 0 && x < n && !seen[x]) { seen[x] = 1; order.push_back(x); } };
    if (getenv("JB200_NO_RENUMBER") == nullptr || atoi(getenv("JB200_NO_RENUMBER")) == 0) {
      for (int i = 0; i < t->n_iso; i++) push(t->iso_node[i]);
      for (int i = 0; i < t->n_shared; i++) push(t->shared_node[i]);
      if (grammar) for (int i = 0; i < t->n_init; i++) push(t->init_node[i]);
      for (size_t q = 0; q < order.size(); q++) {
        const int x = order[q];
        if (t->next_a[x] != JB200_LOG_ZERO) push(x + 1);
        for (int k = t->arc_off[x]; k < t->arc_off[x + 1]; k++) push(t->arc_to[k]);
      }
    }
    for (int x = 0; x < n; x++) push(x);                       // whatever the roots do not reach keeps its relative order
    for (int i = 0; i < n; i++) { perm[order[i]] = i; inv[i] = order[i]; }
  }
  auto remap = [&](const int *src, size_t cnt) { std::vector<int> v(cnt); for (size_t i = 0; i < cnt; i++) v[i] = perm[src[i]]; return v; };
  // node records and the arc lists, in the new order
  std::vector<NodeRec> nodes(n);
  std::vector<int> arc_to_n((size_t)std::max(t->n_arcs, 1)); std::vector<float> arc_a_n((size_t)std::max(t->n_arcs, 1));
  {
    int ao = 0;
    for (int i = 0; i < n; i++) {
      const int o = inv[i];
      NodeRec &r = nodes[i];
      r.self_a = t->self_a[o]; r.next_a = t->next_a[o];
      r.arc_off = ao; r.arc_n = t->arc_off[o + 1] - t->arc_off[o];
      for (int k = t->arc_off[o]; k < t->arc_off[o + 1]; k++) { arc_to_n[ao] = perm[t->arc_to[k]]; arc_a_n[ao] = t->arc_a[k]; ao++; }
      r.stend = t->stend[o]; r.scid = t->scid[o];
      const int style = t->outstyle[o];
      if (style > 3 && !(style == 255 && t->multipath)) { set_error("non-emitting node in a non-multipath tree"); jb200_decoder_destroy(d); return JB200_ERR_UNSUPPORTED; }
      r.out = (style == 255) ? (int)0xF0000000u : (int)(((unsigned)style << 28) | (unsigned)(t->out_ref[o] & 0x0fffffff));
      r.next = (r.next_a != JB200_LOG_ZERO && o + 1 < n) ? perm[o + 1] : i;
    }
  }
  TRY(dev_upload(d, nodes.data(), nodes.size(), &P.nodes));
  TRY(dev_upload(d, arc_to_n.data(), (size_t)t->n_arcs, &P.arc_to));
  TRY(dev_upload(d, arc_a_n.data(), (size_t)t->n_arcs, &P.arc_a));
  TRY(dev_upload(d, t->rset_ctx, (size_t)t->n_rset * (t->n_ctx + 1), &P.rset_ctx));
  TRY(dev_upload(d, t->word_ctx, (size_t)t->n_words, &P.word_ctx));
  P.n_ctx = t->n_ctx;
  { const std::vector<int> v = remap(t->iso_node, (size_t)t->n_iso); TRY(dev_upload(d, v.data(), v.size(), &P.iso_node)); }
  TRY(dev_upload(d, t->iso_id, (size_t)t->n_iso, &P.iso_id));
  P.n_iso = t->n_iso;
  { const std::vector<int> v = remap(t->shared_node, (size_t)t->n_shared); TRY(dev_upload(d, v.data(), v.size(), &P.shared_node)); }
  {
    std::vector<float> sf(std::max(t->n_shared, 1));
    for (int i = 0; i < t->n_shared; i++) {
      const int sc = t->scid[t->shared_node[i]];
      if (sc >= 0) { set_error("shared root without 1-gram factoring value"); jb200_decoder_destroy(d); return JB200_ERR_ARG; }
      sf[i] = t->fscore[-sc];
    }
    TRY(dev_upload(d, sf.data(), (size_t)t->n_shared, &P.shared_f));
  }
  P.n_shared = t->n_shared;
  P.multipath = t->multipath ? 1 : 0;
  P.cp_allowed = nullptr; P.init_node = nullptr; P.init_lscore = nullptr; P.n_init = 0; P.penalty1 = 0.0f;
  if (grammar) {
    TRY(dev_upload(d, t->cp_allowed, (size_t)t->n_words * t->n_iso, &P.cp_allowed));
    { const std::vector<int> v = remap(t->init_node, (size_t)t->n_init); TRY(dev_upload(d, v.data(), v.size(), &P.init_node)); }
    TRY(dev_upload(d, t->init_lscore, (size_t)t->n_init, &P.init_lscore));
    P.n_init = t->n_init; P.penalty1 = t->penalty1;
  }
  {
    // multipath: expand the roots into their successors (self, next, arcs: the order propagation visits them,
    // beam.c:2467-2500); the word-begin node of the head silence is never entered (beam.c:2336-2342)
    std::vector<int> ia_node, ia_iso, sa_node, sa_sh; std::vector<float> ia_a, sa_a;
    if (t->multipath) {
      auto expand = [&](int root, int idx, std::vector<int> &vn, std::vector<int> &vi, std::vector<float> &va) {
        if (t->self_a[root] != JB200_LOG_ZERO) { vn.push_back(perm[root]); vi.push_back(idx); va.push_back(t->self_a[root]); }
        if (t->next_a[root] != JB200_LOG_ZERO) { vn.push_back(perm[root + 1]); vi.push_back(idx); va.push_back(t->next_a[root]); }
        for (int k = t->arc_off[root]; k < t->arc_off[root + 1]; k++) { vn.push_back(perm[t->arc_to[k]]); vi.push_back(idx); va.push_back(t->arc_a[k]); }
      };
      const int head_begin = t->wordbegin[t->head_silwid];
      for (int i = 0; i < t->n_iso; i++) if (t->iso_node[i] != head_begin) expand(t->iso_node[i], i, ia_node, ia_iso, ia_a);
      for (int i = 0; i < t->n_shared; i++) expand(t->shared_node[i], i, sa_node, sa_sh, sa_a);
    }
    P.n_isoarc = (int)ia_node.size(); P.n_sharc = (int)sa_node.size();
    TRY(dev_upload(d, ia_node.data(), ia_node.size(), &P.isoarc_node));
    TRY(dev_upload(d, ia_iso.data(), ia_iso.size(), &P.isoarc_iso));
    TRY(dev_upload(d, ia_a.data(), ia_a.size(), &P.isoarc_a));
    TRY(dev_upload(d, sa_node.data(), sa_node.size(), &P.sharc_node));
    TRY(dev_upload(d, sa_sh.data(), sa_sh.size(), &P.sharc_shared));
    TRY(dev_upload(d, sa_a.data(), sa_a.size(), &P.sharc_a));
    if (t->n_iso + 1024 >= (int)SEQ_LOCAL || P.n_isoarc >= (int)SEQ_LOCAL || P.n_sharc >= (int)SEQ_LOCAL || t->n_shared >= (int)SEQ_LOCAL) {
      set_error("too many tree roots (%d isolated, %d shared) for the arrival-order numbering", t->n_iso, t->n_shared);
      jb200_decoder_destroy(d); return JB200_ERR_UNSUPPORTED;
    }
  }
  TRY(dev_upload(d, t->wordend_a, (size_t)t->n_words, &P.wordend_a));
  TRY(dev_upload(d, t->is_transparent, (size_t)t->n_words, &P.is_transp));
  TRY(dev_upload(d, t->wton, (size_t)t->n_words, &P.wton));
  TRY(dev_upload(d, t->cprob, (size_t)t->n_words, &P.cprob));
  TRY(dev_upload(d, t->fscore, (size_t)t->n_fscore, &P.fscore));
  TRY(dev_upload(d, t->scword, (size_t)t->n_scword, &P.scword));
  TRY(dev_upload(d, t->uni_prob, (size_t)t->lm_nvocab, &P.uni_prob));
  TRY(dev_upload(d, t->uni_bow, (size_t)t->lm_nvocab, &P.uni_bow));
  TRY(dev_upload(d, t->bi_bgn, (size_t)t->lm_nvocab, &P.bi_bgn));
  TRY(dev_upload(d, t->bi_num, (size_t)t->lm_nvocab, &P.bi_num));
  TRY(dev_upload(d, t->bi_wid, (size_t)t->lm_nbigram, &P.bi_wid));
  TRY(dev_upload(d, t->bi_prob, (size_t)t->lm_nbigram, &P.bi_prob));
  P.lm_mode = t->lm_mode; P.lm_unk_id = t->lm_unk_id; P.lm_unk_num_log = t->lm_unk_num_log;
  P.lm_weight = t->lm_weight; P.lm_penalty = t->lm_penalty; P.lm_penalty_trans = t->lm_penalty_trans;
  P.prune_width = t->score_pruning_width;
  P.head_node = grammar ? 0 : perm[t->wordbegin[t->head_silwid]]; P.tail_silwid = t->tail_silwid; P.beam = t->beam_width; P.n_nodes = n;
  // cd sets come from the AM handle's descriptor: re-upload from the gmm handle is not exposed, so the
  // decoder asks the scorer for its device copies
  {
    const int *co = nullptr, *cs = nullptr; int meth = 0, nb = 0;
    gmm_cd_device(am, &co, &cs, &meth, &nb);
    P.cd_off = co; P.cd_states = cs; P.iwcd_method = meth; P.iwcd_nbest = nb;
  }
  // inter-word bigram rows for every last word (the reference's iw_sc_cache, fully populated)
  {
    const int *d_iso_word = nullptr;
    TRY(dev_upload(d, t->iso_word, (size_t)t->n_iso, &d_iso_word));
    float *iw = nullptr;
    TRY(dev_alloc_shared(d, (size_t)t->n_words * std::max(t->n_iso, 1), &iw));
    P.iw = iw;
    const long long tot = grammar ? 0 : (long long)t->n_words * t->n_iso;     // grammar mode reads cp_allowed instead
    if (tot > 0) {
      iw_table_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, d->stream>>>(P, d_iso_word, iw, t->n_words);
      g_launches.fetch_add(1);
      TRYC(cudaGetLastError());
    }
  }
  // work areas
  // the reference starts at 2*beam+startnum tokens and grows on demand; we size once and flag overflow.  Seen on the
  // 20k-word tree: 4.6*beam at -b 800 (of which startnum = 1375 root tokens), 4.7*beam at -b 4000.  The array lives
  // in shared memory, and what it takes is lost to L1 (5*beam+startnum at -b 800 costs 30 % of the kernel's speed).
  int maxt = (std::max(4 * t->beam_width + t->n_start, 5 * t->beam_width) + 64 + 3) & ~3;
  if (const char *e = getenv("JB200_MAXT")) maxt = (std::max(atoi(e), 64) + 3) & ~3;
  // Where the heap-select array lives.  Shared memory as long as one block's share fits; a wide beam on a large tree
  // (-b 4000 on the 60k-word multipath tree creates up to 8.5 x beam tokens a frame) goes to global memory instead,
  // with room for 9 x beam + startnum tokens, and shared memory keeps only the closed form's sort area.
  const size_t offs_bytes = (size_t)(t->beam_width + 2) * 4 * 2;
  int smem_limit = 0;
  TRYC(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, d->device));
  smem_limit -= 2048;                                       // static shared variables of the kernels
  bool heap_global = (size_t)(maxt + 4) * 8 + offs_bytes > (size_t)smem_limit / 2;   // would leave one block per SM
  if (const char *e = getenv("JB200_HEAP_GLOBAL")) heap_global = atoi(e) != 0;
  P.heap_g = nullptr; P.sort_cap = 0; P.qcap = 0;
  if (heap_global) {
    if (!getenv("JB200_MAXT")) maxt = std::min(65000, (std::max(maxt, 9 * t->beam_width + t->n_start) + 3) & ~3);
    int sc = 1024; while (sc < 2 * t->beam_width && sc < 16384) sc <<= 1;       // candidates = beam + one histogram bin
    const size_t tail_bytes = (size_t)(t->beam_width + 2) * 8;                   // the replay's copy of the tail slots
    if ((size_t)sc * 8 + offs_bytes + tail_bytes > (size_t)smem_limit) { set_error("beam width %d needs more shared memory than the device has", t->beam_width); jb200_decoder_destroy(d); return JB200_ERR_UNSUPPORTED; }
    P.sort_cap = sc;
    // what is left of shared memory holds the top levels of the heap during a replay (one block per SM: the replay is all
    // that matters at these beam widths); JB200_HEAP_CACHE=0 keeps only the sort area (two blocks per SM)
    int qc = sc;
    if (!getenv("JB200_HEAP_CACHE") || atoi(getenv("JB200_HEAP_CACHE")) != 0)
      while ((size_t)qc * 2 * 8 + offs_bytes + tail_bytes <= (size_t)smem_limit && qc * 2 <= ((maxt + 4) | 1023) + 1) qc <<= 1;
    P.qcap = qc;
    TRY(dev_alloc(d, (size_t)max_utts * (maxt + 4), &P.heap_g));
  }
  P.maxt = maxt; P.maxc = 4 * maxt; P.maxw = t->beam_width + 1;
  TRY(dev_alloc(d, (size_t)max_utts * 2 * maxt, &P.tok));
  TRY(dev_alloc(d, (size_t)max_utts * 2 * maxt, &P.order));
  TRY(dev_alloc(d, (size_t)max_utts * n, &P.slots));
  TRY(dev_alloc(d, (size_t)max_utts * P.maxc, &P.cand));
  TRY(dev_alloc(d, (size_t)max_utts * P.maxc, &P.candb));
  TRY(dev_alloc(d, (size_t)max_utts * (t->beam_width + 2), &P.surv));
  const int n_isoent = std::max(std::max(t->n_iso, P.n_isoarc), 1);
  TRY(dev_alloc(d, (size_t)max_utts * n_isoent, &P.iso));
  TRY(dev_alloc(d, (size_t)max_utts * P.maxw, &P.wend));
  P.maxbits = (P.maxc + std::min(P.maxw, 256) * n_isoent + std::max(t->n_shared, P.n_sharc) + 63) & ~31;
  TRY(dev_alloc(d, (size_t)max_utts * (P.maxbits >> 5), &P.bitmask));
  TRY(dev_alloc(d, (size_t)max_utts * (P.maxbits >> 5), &P.wordpre));
  TRY(dev_alloc(d, 8, &P.misspec_counter));        // [0] mis-speculations, [1] replay ticks (levels), [2] extractions, [3] held-back starts, [4] upward selects, [5] of which closed form
  TRYC(cudaMemset(P.misspec_counter, 0, 8 * sizeof(unsigned long long)));
  P.force_seq_heap = getenv("JB200_FORCE_SEQ_HEAP") ? atoi(getenv("JB200_FORCE_SEQ_HEAP")) : 0;
  P.check_heap = getenv("JB200_CHECK_HEAP") ? atoi(getenv("JB200_CHECK_HEAP")) : 0;
  {
    // bigram-factoring memo: keys are (word id, successor slot) packed 16+16, so it needs both below 65535
    int bits = getenv("JB200_LMCACHE_BITS") ? atoi(getenv("JB200_LMCACHE_BITS")) : 21;
    if (t->n_words >= 65535 || t->n_scword >= 65535 || bits < 8) bits = 0;
    if (bits > 26) bits = 26;
    P.lmc_bits = bits; P.lmc = nullptr;
    if (bits > 0) {
      TRY(dev_alloc_shared(d, (size_t)1 << bits, &P.lmc));
      TRYC(cudaMemsetAsync(P.lmc, 0xff, sizeof(unsigned long long) << bits, d->stream));
    }
  }
  P.chunk = nullptr; P.state = nullptr; P.interim = 0; P.interim_words = nullptr; P.atoms_in_place = 0;
  P.no_reloc = getenv("JB200_NO_RELOCATE") ? atoi(getenv("JB200_NO_RELOCATE")) : 0;
  P.prof_fine = getenv("JB200_PROF_FINE") ? atoi(getenv("JB200_PROF_FINE")) : 0;   // extra barrier: slot 4 = word-internal expansion alone
  P.no_lose = getenv("JB200_NO_LOSER_CUT") ? atoi(getenv("JB200_NO_LOSER_CUT")) : 0;
  P.no_closed = getenv("JB200_NO_CLOSED_FORM") ? atoi(getenv("JB200_NO_CLOSED_FORM")) : 0;   // 1: always replay the extraction loop
  P.heap_single = getenv("JB200_HEAP_SINGLE") ? atoi(getenv("JB200_HEAP_SINGLE")) : 0;   // 1: the single-thread replay, 2: the readable pipelined loop, 3: the C++ form of the shipped loop (A/B timing)
  {
    size_t tot = (size_t)max_utts * n;
    fill_slots_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, d->stream>>>(P.slots, tot);
    g_launches.fetch_add(1);
    TRYC(cudaGetLastError());
  }
  d->atoms_cap = (long long)max_frames * d->atoms_per_frame + (long long)max_utts * 64;
  TRY(dev_alloc(d, (size_t)d->atoms_cap, &P.atoms_raw));
  TRY(dev_alloc(d, (size_t)d->atoms_cap, &P.newidx));
  TRY(dev_alloc(d, (size_t)max_frames + max_utts + 8, &P.group0));
  TRY(dev_alloc(d, (size_t)max_frames * 2 + 8, &P.counts));
  TRY(dev_alloc(d, (size_t)d->atoms_cap, &d->d_atoms_out));
  TRY(dev_alloc(d, 1, &d->d_atom_counter));
  TRY(dev_alloc(d, (size_t)max_utts, &d->d_results));
  TRY(dev_alloc(d, (size_t)max_utts * MAX_WORDS, &d->d_words));
  TRY(dev_alloc(d, (size_t)max_utts * 8, &d->d_prof));
  TRY(dev_alloc(d, (size_t)max_utts + 1, &d->d_frame_off));
  TRY(dev_alloc(d, (size_t)max_utts + 1, &d->d_atom_off));
  TRY(dev_alloc(d, (size_t)max_frames * d->dim, &d->d_feats));
  TRY(dev_alloc(d, (size_t)max_frames * d->row_stride, &d->d_rows));
  TRYC(cudaMallocHost(&d->h_results, sizeof(jb200_utt_result) * max_utts));
  TRYC(cudaMallocHost(&d->h_atoms, sizeof(jb200_atom) * (size_t)d->atoms_cap));
  TRYC(cudaMallocHost(&d->h_words, sizeof(int) * (size_t)max_utts * MAX_WORDS));
  TRYC(cudaMallocHost(&d->h_counter, sizeof(unsigned long long)));
  TRY(dev_alloc(d, (size_t)jb200_decoder::MAX_CHUNKS * max_utts, &d->d_chunk));
  TRYC(cudaMallocHost(&d->h_chunk, sizeof(ChunkDesc) * (size_t)jb200_decoder::MAX_CHUNKS * max_utts));
  TRY(dev_alloc(d, (size_t)max_utts, &d->d_state));
  TRYC(cudaMemset(d->d_state, 0, sizeof(UttState) * (size_t)max_utts));
  TRY(dev_alloc(d, (size_t)max_utts * MAX_WORDS, &d->d_interim_words));
  TRY(dev_alloc(d, (size_t)jb200_decoder::MAX_CHUNKS * (2 * max_utts + 1), &d->d_seg));
  TRYC(cudaMallocHost(&d->h_seg, sizeof(int) * (size_t)jb200_decoder::MAX_CHUNKS * (2 * max_utts + 1)));
  TRYC(cudaMallocHost(&d->h_aoff, sizeof(long long) * (size_t)(max_utts + 1)));
  TRYC(cudaMallocHost(&d->h_state, sizeof(UttState) * (size_t)max_utts));
  TRYC(cudaMallocHost(&d->h_interim_words, sizeof(int) * (size_t)max_utts * MAX_WORDS));
  {
    // the scoring stream of the batch pipeline gets the higher priority: its thread blocks take the room the token-passing
    // kernel leaves on every SM as soon as it is free
    int lo = 0, hi = 0;
    TRYC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    TRYC(cudaStreamCreateWithPriority(&d->score_stream, cudaStreamNonBlocking, hi));
    for (auto &e : d->ev_slice) TRYC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    TRYC(cudaEventCreate(&d->ev_score_begin)); TRYC(cudaEventCreate(&d->ev_score_end));
    d->pipe_frames = 0;
    if (const char *e = getenv("JB200_PIPE_FRAMES")) d->pipe_frames = std::max(0, atoi(e));
  }
  d->smem_bytes = (heap_global ? (size_t)P.qcap * 8 + (size_t)(t->beam_width + 2) * 8 : (size_t)(maxt + 4) * 8) + offs_bytes;
  d->grammar = grammar;
  const void *kern = grammar ? (const void *)beam_kernel_grammar : P.multipath ? (const void *)beam_kernel_mp : (const void *)beam_kernel;
  TRYC(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_bytes));
  {
    int per_sm = 0, sms = 0;
    TRYC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BEAM_THREADS, d->smem_bytes));
    TRYC(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d->device));
    d->resident = per_sm * sms;
  }
  // Opt-in (JB200_L2_WINDOW=1): an L2 access-policy window that keeps the shared tables resident (persisting hits, streaming
  // misses).  Measured on the 20k-word workload with 592 distinct utterances: 1.194 M frames/s with the window, 1.224 M
  // without -- the set-aside costs the per-utterance work areas more than it saves on tree / LM lines -- so it is off.
  if (d->arena && d->arena_used > 0 && getenv("JB200_L2_WINDOW") != nullptr && atoi(getenv("JB200_L2_WINDOW")) != 0) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
      const size_t win = std::min<size_t>(d->arena_used, (size_t)prop.accessPolicyMaxWindowSize);
      const size_t carve = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, win);
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr = d->arena;
        av.accessPolicyWindow.num_bytes = win;
        av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)win);
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        d->l2_window = (cudaStreamSetAttribute(d->stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess);
      }
    }
    cudaGetLastError();
  }
  TRYC(cudaStreamSynchronize(d->stream));
#undef TRY
#undef TRYC
  *out = d;
  return JB200_OK;
}

// Cut a batch into time slices.  One slice (the whole utterance per launch) unless the pipeline is on: then slice c holds
// frames [c*F, (c+1)*F) of every utterance, scored by its own launch on the scoring stream (a gather over the segment
// list) while the beam kernel works on slice c-1.  Fills the staging copies of the chunk descriptors and segment lists.
static int plan_slices(jb200_decoder *d, const int32_t *frame_off, int n_utts, bool allow_pipe) {
  int maxT = 0;
  for (int u = 0; u < n_utts; u++) maxT = std::max(maxT, frame_off[u + 1] - frame_off[u]);
  int F = (allow_pipe && !d->dnn && d->pipe_frames > 0) ? d->pipe_frames : 0;
  int nch = 1;
  if (F > 0) {
    nch = (maxT + F - 1) / F;
    if (nch > jb200_decoder::MAX_CHUNKS) { F = (maxT + jb200_decoder::MAX_CHUNKS - 1) / jb200_decoder::MAX_CHUNKS; nch = (maxT + F - 1) / F; }
    if (nch < 2) { nch = 1; F = 0; }
  }
  d->n_chunks = nch; d->last_piped = (F > 0);
  d->slice_nseg.assign(nch, 0); d->slice_frames.assign(nch, 0);
  const int mu = d->max_utts, segw = 2 * mu + 1;
  for (int c = 0; c < nch; c++) {
    ChunkDesc *cd = d->h_chunk + (size_t)c * mu;
    int *so = d->h_seg + (size_t)c * segw, *ss = so + mu + 1;     // seg_off [mu+1], seg_start [mu]
    int nseg = 0, lf = 0;
    for (int u = 0; u < n_utts; u++) {
      const int T = frame_off[u + 1] - frame_off[u];
      ChunkDesc k;
      k.row_base = frame_off[u];
      if (F == 0) { k.t0 = 0; k.t1 = T; k.flags = CHUNK_FIRST | CHUNK_FINAL; }
      else {
        k.t0 = std::min(c * F, T); k.t1 = std::min((c + 1) * F, T);
        const bool past = (c > 0) && (c * F >= T);             // the utterance ended in an earlier slice
        const bool last = ((c + 1) * F >= T);
        k.flags = past ? CHUNK_SKIP : ((c == 0 ? CHUNK_FIRST : 0) | (last ? CHUNK_FINAL : 0));
        if (!past && k.t1 > k.t0) { so[nseg] = lf; ss[nseg] = frame_off[u] + k.t0; nseg++; lf += k.t1 - k.t0; }
      }
      cd[u] = k;
    }
    so[nseg] = lf;
    d->slice_nseg[c] = nseg; d->slice_frames[c] = lf;
  }
  return JB200_OK;
}

static int prepare_batch(jb200_decoder *d, const int32_t *frame_off, int n_utts, bool allow_pipe = true) {
  if (!d || !frame_off || n_utts < 1) { set_error("decode: bad argument"); return JB200_ERR_ARG; }
  if (n_utts > d->max_utts) { set_error("batch of %d utterances exceeds decoder capacity %d", n_utts, d->max_utts); return JB200_ERR_CAPACITY; }
  const int total = frame_off[n_utts] - frame_off[0];
  if (frame_off[0] != 0) { set_error("frame_off[0] must be 0"); return JB200_ERR_ARG; }
  if (total > d->max_frames) { set_error("batch of %d frames exceeds decoder capacity %d", total, d->max_frames); return JB200_ERR_CAPACITY; }
  JB_CUDA(cudaSetDevice(d->device));
  JB_CUDA(cudaStreamSynchronize(d->stream));   // the staging buffers below may still feed the previous batch's copies
  d->h_aoff[0] = 0;
  for (int u = 0; u < n_utts; u++) {
    const int T = frame_off[u + 1] - frame_off[u];
    if (T < 0 || T > 32767) { set_error("utterance %d has %d frames (trellis times are 16-bit in the reference)", u, T); return JB200_ERR_ARG; }
    d->h_aoff[u + 1] = d->h_aoff[u] + (long long)T * d->atoms_per_frame + 64;
  }
  d->stream_mode = false;
  d->h_frame_off.assign(frame_off, frame_off + n_utts + 1);
  int rc = plan_slices(d, frame_off, n_utts, allow_pipe); if (rc) return rc;
  const int mu = d->max_utts;
  JB_CUDA(cudaMemcpyAsync(d->d_frame_off, d->h_frame_off.data(), sizeof(int) * (n_utts + 1), cudaMemcpyHostToDevice, d->stream));
  JB_CUDA(cudaMemcpyAsync(d->d_atom_off, d->h_aoff, sizeof(long long) * (n_utts + 1), cudaMemcpyHostToDevice, d->stream));
  JB_CUDA(cudaMemcpyAsync(d->d_chunk, d->h_chunk, sizeof(ChunkDesc) * (size_t)d->n_chunks * mu, cudaMemcpyHostToDevice, d->stream));
  if (d->last_piped)
    JB_CUDA(cudaMemcpyAsync(d->d_seg, d->h_seg, sizeof(int) * (size_t)d->n_chunks * (2 * mu + 1), cudaMemcpyHostToDevice, d->stream));
  JB_CUDA(cudaMemsetAsync(d->d_atom_counter, 0, sizeof(unsigned long long), d->stream));
  JB_CUDA(cudaStreamSynchronize(d->stream));   // h_frame_off is a std::vector (pageable)
  d->last_n = n_utts; d->last_total_frames = total; d->fetched = false;
  return JB200_OK;
}

static int launch_beam(jb200_decoder *d, int n_utts, int chunk_index = 0, int interim = 0) {
  BeamParams P = d->P;
  P.rows = d->d_rows; P.row_stride = d->row_stride; P.frame_off = d->d_frame_off;
  P.atom_off = d->d_atom_off; P.atoms_out = d->d_atoms_out; P.atom_counter = d->d_atom_counter;
  P.atoms_out_cap = d->atoms_cap; P.results = d->d_results; P.words = d->d_words; P.prof = d->d_prof;
  P.chunk = d->d_chunk + (size_t)chunk_index * d->max_utts; P.state = d->d_state;
  P.interim = interim; P.interim_words = d->d_interim_words; P.atoms_in_place = d->stream_mode ? 1 : 0;
  if (d->grammar) beam_kernel_grammar<<<n_utts, BEAM_THREADS, d->smem_bytes, d->stream>>>(P);
  else if (P.multipath) beam_kernel_mp<<<n_utts, BEAM_THREADS, d->smem_bytes, d->stream>>>(P);
  else beam_kernel<<<n_utts, BEAM_THREADS, d->smem_bytes, d->stream>>>(P);
  JB_LAUNCH_CHECK();
  return JB200_OK;
}

// scoring + token passing of a prepared batch whose features are at d_feats; ev[1] has been recorded on the main stream
static int run_batch(jb200_decoder *d, const float *d_feats, int n_utts) {
  int rc;
  if (!d->last_piped) {
    rc = d->dnn ? dnn_forward_device(d->dnn, d_feats, d->last_total_frames, d->d_rows, d->row_stride, d->stream)
                : gmm_launch_states(d->am, d_feats, d->last_total_frames, d->d_rows, d->row_stride, d->stream, nullptr, nullptr, 0);
    if (rc) return rc;
    JB_CUDA(cudaEventRecord(d->ev[2], d->stream));
    return launch_beam(d, n_utts, 0);
  }
  // pipeline: every slice's scoring is queued on the scoring stream at once (it only depends on the features), the beam
  // kernel of slice c waits for the scores of slice c alone
  const int mu = d->max_utts, segw = 2 * mu + 1;
  JB_CUDA(cudaStreamWaitEvent(d->score_stream, d->ev[1], 0));
  JB_CUDA(cudaEventRecord(d->ev_score_begin, d->score_stream));
  for (int c = 0; c < d->n_chunks; c++) {
    const int *ds = d->d_seg + (size_t)c * segw;
    rc = gmm_launch_states(d->am, d_feats, d->slice_frames[c], d->d_rows, d->row_stride, d->score_stream, ds, ds + mu + 1, d->slice_nseg[c]);
    if (rc) return rc;
    JB_CUDA(cudaEventRecord(d->ev_slice[c], d->score_stream));
  }
  JB_CUDA(cudaEventRecord(d->ev_score_end, d->score_stream));
  for (int c = 0; c < d->n_chunks; c++) {
    JB_CUDA(cudaStreamWaitEvent(d->stream, d->ev_slice[c], 0));
    if (c == 0) JB_CUDA(cudaEventRecord(d->ev[2], d->stream));       // "scoring" = what the beam had to wait for
    rc = launch_beam(d, n_utts, c); if (rc) return rc;
  }
  return JB200_OK;
}

extern "C" int jb200_decoder_fetch(jb200_decoder *d) {
  if (!d) { set_error("null decoder"); return JB200_ERR_ARG; }
  if (d->fetched) return JB200_OK;
  JB_CUDA(cudaSetDevice(d->device));
  JB_CUDA(cudaEventRecord(d->ev[3], d->stream));
  JB_CUDA(cudaMemcpyAsync(d->h_counter, d->d_atom_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, d->stream));
  JB_CUDA(cudaMemcpyAsync(d->h_results, d->d_results, sizeof(jb200_utt_result) * d->last_n, cudaMemcpyDeviceToHost, d->stream));
  JB_CUDA(cudaMemcpyAsync(d->h_words, d->d_words, sizeof(int) * (size_t)d->last_n * MAX_WORDS, cudaMemcpyDeviceToHost, d->stream));
  JB_CUDA(cudaStreamSynchronize(d->stream));
  long long na = (long long)*d->h_counter;
  if (na > d->atoms_cap) na = d->atoms_cap;
  d->last_atoms = na;
  if (na > 0) JB_CUDA(cudaMemcpyAsync(d->h_atoms, d->d_atoms_out, sizeof(jb200_atom) * (size_t)na, cudaMemcpyDeviceToHost, d->stream));
  JB_CUDA(cudaEventRecord(d->ev[4], d->stream));
  JB_CUDA(cudaStreamSynchronize(d->stream));
  cudaEventElapsedTime(&d->last_ms[0], d->ev[0], d->ev[1]);
  cudaEventElapsedTime(&d->last_ms[1], d->ev[1], d->ev[2]);
  cudaEventElapsedTime(&d->last_ms[2], d->ev[2], d->ev[3]);
  cudaEventElapsedTime(&d->last_ms[3], d->ev[3], d->ev[4]);
  if (d->last_piped) cudaEventElapsedTime(&d->last_score_busy_ms, d->ev_score_begin, d->ev_score_end);
  d->last_d2h = (long long)sizeof(unsigned long long) + (long long)sizeof(jb200_utt_result) * d->last_n +
                (long long)sizeof(int) * d->last_n * MAX_WORDS + (long long)sizeof(jb200_atom) * na;
  d->fetched = true;
  return JB200_OK;
}

extern "C" int jb200_decoder_sync_timing(jb200_decoder *d) {
  if (!d) { set_error("null decoder"); return JB200_ERR_ARG; }
  JB_CUDA(cudaSetDevice(d->device));
  JB_CUDA(cudaEventSynchronize(d->ev[3]));
  cudaEventElapsedTime(&d->last_ms[0], d->ev[0], d->ev[1]);
  cudaEventElapsedTime(&d->last_ms[1], d->ev[1], d->ev[2]);
  cudaEventElapsedTime(&d->last_ms[2], d->ev[2], d->ev[3]);
  if (d->last_piped) cudaEventElapsedTime(&d->last_score_busy_ms, d->ev_score_begin, d->ev_score_end);
  d->last_ms[3] = 0.0f;
  return JB200_OK;
}

extern "C" int jb200_decoder_phase_cycles(jb200_decoder *d, int64_t *cycles, int n_utts) {
  if (!d || !cycles || n_utts < 1 || n_utts > d->last_n) { set_error("bad argument"); return JB200_ERR_ARG; }
  JB_CUDA(cudaSetDevice(d->device));
  JB_CUDA(cudaMemcpy(cycles, d->d_prof, sizeof(long long) * 8 * (size_t)n_utts, cudaMemcpyDeviceToHost));
  return JB200_OK;
}

extern "C" int jb200_dnn_in_dim(const jb200_dnn *h);
extern "C" int jb200_dnn_out_dim(const jb200_dnn *h);
extern "C" int jb200_decoder_attach_dnn(jb200_decoder *d, jb200_dnn *dnn) {
  if (!d || !dnn) { set_error("null argument"); return JB200_ERR_ARG; }
  if (jb200_dnn_out_dim(dnn) != d->S) { set_error("DNN has %d outputs but the HMM set has %d states", jb200_dnn_out_dim(dnn), d->S); return JB200_ERR_ARG; }
  JB_CUDA(cudaSetDevice(d->device));
  const int dim = jb200_dnn_in_dim(dnn);
  if (dim != d->dim) {
    // the feature buffer was sized for the AM's dimension; re-size it for the DNN's input width
    float *nf = nullptr;
    JB_CUDA(cudaMalloc(&nf, sizeof(float) * (size_t)d->max_frames * dim));
    d->dev_allocs.push_back(nf);
    d->d_feats = nf; d->dim = dim;
  }
  d->dnn = dnn;
  return JB200_OK;
}

extern "C" int64_t jb200_decoder_misspeculations(jb200_decoder *d) {
  if (!d) return -1;
  unsigned long long v = 0;
  cudaSetDevice(d->device);
  if (cudaMemcpy(&v, d->P.misspec_counter, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)v;
}

extern "C" int jb200_decoder_heap_stats(jb200_decoder *d, int64_t out[3]) {
  if (!d || !out) { set_error("bad argument"); return JB200_ERR_ARG; }
  unsigned long long v[3] = {0, 0, 0};
  JB_CUDA(cudaSetDevice(d->device));
  JB_CUDA(cudaMemcpy(v, d->P.misspec_counter, sizeof(v), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 3; i++) out[i] = (int64_t)v[i];
  return JB200_OK;
}

extern "C" int jb200_decoder_select_stats(jb200_decoder *d, int64_t out[2]) {
  if (!d || !out) { set_error("bad argument"); return JB200_ERR_ARG; }
  unsigned long long v[2] = {0, 0};
  JB_CUDA(cudaSetDevice(d->device));
  JB_CUDA(cudaMemcpy(v, d->P.misspec_counter + 4, sizeof(v), cudaMemcpyDeviceToHost));
  out[0] = (int64_t)v[0]; out[1] = (int64_t)v[1];
  return JB200_OK;
}

extern "C" int64_t jb200_decoder_relocated_selects(jb200_decoder *d) {
  if (!d) return -1;
  unsigned long long v = 0;
  cudaSetDevice(d->device);
  if (cudaMemcpy(&v, d->P.misspec_counter + 6, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)v;
}

extern "C" int64_t jb200_decoder_last_d2h_bytes(const jb200_decoder *d) { return d ? d->last_d2h : 0; }
extern "C" int jb200_decoder_resident_utts(const jb200_decoder *d) { return d ? d->resident : 0; }

extern "C" int jb200_decode_batch_device(jb200_decoder *d, const float *d_feats, const int32_t *frame_off, int n_utts) {
  int rc = prepare_batch(d, frame_off, n_utts); if (rc) return rc;
  JB_CUDA(cudaEventRecord(d->ev[0], d->stream));
  JB_CUDA(cudaEventRecord(d->ev[1], d->stream));
  rc = run_batch(d, d_feats, n_utts); if (rc) return rc;
  JB_CUDA(cudaEventRecord(d->ev[3], d->stream));
  return JB200_OK;
}

extern "C" int jb200_decode_batch_host(jb200_decoder *d, const float *feats, const int32_t *frame_off, int n_utts) {
  if (!feats) { set_error("null feats"); return JB200_ERR_ARG; }
  int rc = prepare_batch(d, frame_off, n_utts); if (rc) return rc;
  JB_CUDA(cudaEventRecord(d->ev[0], d->stream));
  JB_CUDA(cudaMemcpyAsync(d->d_feats, feats, sizeof(float) * (size_t)d->last_total_frames * d->dim, cudaMemcpyHostToDevice, d->stream));
  JB_CUDA(cudaEventRecord(d->ev[1], d->stream));
  rc = run_batch(d, d->d_feats, n_utts); if (rc) return rc;
  return jb200_decoder_fetch(d);
}

extern "C" int jb200_decode_batch_scores_host(jb200_decoder *d, const float *scores, const int32_t *frame_off, int n_utts) {
  if (!scores) { set_error("null scores"); return JB200_ERR_ARG; }
  int rc = prepare_batch(d, frame_off, n_utts, false); if (rc) return rc;
  JB_CUDA(cudaEventRecord(d->ev[0], d->stream));
  JB_CUDA(cudaMemcpy2DAsync(d->d_rows, sizeof(float) * d->row_stride, scores, sizeof(float) * d->S, sizeof(float) * d->S,
                            d->last_total_frames, cudaMemcpyHostToDevice, d->stream));
  JB_CUDA(cudaEventRecord(d->ev[1], d->stream));
  JB_CUDA(cudaEventRecord(d->ev[2], d->stream));
  rc = launch_beam(d, n_utts, 0); if (rc) return rc;
  return jb200_decoder_fetch(d);
}

// ---- frame-synchronous operation (streams) -----------------------------------------------------------------------
extern "C" int jb200_stream_open(jb200_decoder *d, int n_streams) {
  if (!d || n_streams < 1) { set_error("jb200_stream_open: bad argument"); return JB200_ERR_ARG; }
  if (n_streams > d->max_utts) { set_error("%d streams exceed decoder capacity %d", n_streams, d->max_utts); return JB200_ERR_CAPACITY; }
  JB_CUDA(cudaSetDevice(d->device));
  JB_CUDA(cudaStreamSynchronize(d->stream));
  // a stream that was abandoned before its last frame has left node slots behind: wipe those work areas
  if (d->stream_mode) {
    for (int u = 0; u < d->st_n; u++) if (d->st_started[u] && !d->st_done[u]) {
      const size_t tot = (size_t)d->P.n_nodes;
      fill_slots_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, d->stream>>>(d->P.slots + (size_t)u * d->P.n_nodes, tot);
      JB_LAUNCH_CHECK();
    }
  }
  const int cap = std::min(d->max_frames / n_streams, 32767);
  if (cap < 1) { set_error("decoder capacity of %d frames is too small for %d streams", d->max_frames, n_streams); return JB200_ERR_CAPACITY; }
  d->stream_mode = true; d->st_n = n_streams; d->st_cap = cap;
  d->st_t.assign(n_streams, 0); d->st_started.assign(n_streams, 0); d->st_done.assign(n_streams, 0);
  d->h_frame_off.assign(n_streams + 1, 0);
  d->h_aoff[0] = 0;
  for (int u = 0; u < n_streams; u++) {
    d->h_frame_off[u + 1] = d->h_frame_off[u] + cap;
    d->h_aoff[u + 1] = d->h_aoff[u] + (long long)cap * d->atoms_per_frame + 64;
  }
  JB_CUDA(cudaMemcpyAsync(d->d_frame_off, d->h_frame_off.data(), sizeof(int) * (n_streams + 1), cudaMemcpyHostToDevice, d->stream));
  JB_CUDA(cudaMemcpyAsync(d->d_atom_off, d->h_aoff, sizeof(long long) * (n_streams + 1), cudaMemcpyHostToDevice, d->stream));
  JB_CUDA(cudaStreamSynchronize(d->stream));
  d->last_n = n_streams; d->n_chunks = 1; d->last_piped = false; d->fetched = true;
  return JB200_OK;
}

// device part shared by the feature / score variants: rows of the new frames are in d_rows, packed stream-major
static int stream_advance(jb200_decoder *d, const int32_t *n_new, const uint8_t *last, int want_interim) {
  int pack = 0, any = 0;
  for (int u = 0; u < d->st_n; u++) {
    ChunkDesc k; k.t0 = d->st_t[u]; k.t1 = k.t0 + n_new[u]; k.row_base = pack - k.t0; k.flags = 0;
    const bool fin = last && last[u];
    if (d->st_done[u] || (n_new[u] == 0 && !fin)) k.flags = CHUNK_SKIP;
    else {
      if (!d->st_started[u]) k.flags |= CHUNK_FIRST;
      if (fin) k.flags |= CHUNK_FINAL;
      any = 1;
    }
    d->h_chunk[u] = k;
    pack += n_new[u];
  }
  if (!any) return JB200_OK;
  JB_CUDA(cudaMemcpyAsync(d->d_chunk, d->h_chunk, sizeof(ChunkDesc) * (size_t)d->st_n, cudaMemcpyHostToDevice, d->stream));
  int rc = launch_beam(d, d->st_n, 0, want_interim); if (rc) return rc;
  // results of the streams that ended; interim state of the others
  bool fin_any = false;
  for (int u = 0; u < d->st_n; u++) if (d->h_chunk[u].flags & CHUNK_FINAL) fin_any = true;
  if (fin_any) {
    JB_CUDA(cudaMemcpyAsync(d->h_results, d->d_results, sizeof(jb200_utt_result) * d->st_n, cudaMemcpyDeviceToHost, d->stream));
    JB_CUDA(cudaMemcpyAsync(d->h_words, d->d_words, sizeof(int) * (size_t)d->st_n * MAX_WORDS, cudaMemcpyDeviceToHost, d->stream));
  }
  JB_CUDA(cudaMemcpyAsync(d->h_state, d->d_state, sizeof(UttState) * (size_t)d->st_n, cudaMemcpyDeviceToHost, d->stream));
  if (want_interim)
    JB_CUDA(cudaMemcpyAsync(d->h_interim_words, d->d_interim_words, sizeof(int) * (size_t)d->st_n * MAX_WORDS, cudaMemcpyDeviceToHost, d->stream));
  JB_CUDA(cudaStreamSynchronize(d->stream));
  d->last_d2h = 0;
  for (int u = 0; u < d->st_n; u++) {
    const int fl = d->h_chunk[u].flags;
    if (fl & CHUNK_SKIP) continue;
    d->st_started[u] = 1; d->st_t[u] = d->h_chunk[u].t1;
    if (fl & CHUNK_FINAL) {
      d->st_done[u] = 1;
      const int na = d->h_results[u].n_atoms;
      if (na > 0) JB_CUDA(cudaMemcpyAsync(d->h_atoms + d->h_aoff[u], d->d_atoms_out + d->h_aoff[u], sizeof(jb200_atom) * (size_t)na, cudaMemcpyDeviceToHost, d->stream));
      d->last_d2h += (long long)sizeof(jb200_atom) * na + (long long)sizeof(jb200_utt_result) + (long long)sizeof(int) * MAX_WORDS;
    }
  }
  JB_CUDA(cudaStreamSynchronize(d->stream));
  return JB200_OK;
}

static int stream_check(jb200_decoder *d, const int32_t *n_new, int *total) {
  if (!d || !n_new) { set_error("jb200_stream_feed: bad argument"); return JB200_ERR_ARG; }
  if (!d->stream_mode) { set_error("jb200_stream_feed: call jb200_stream_open first"); return JB200_ERR_ARG; }
  int tot = 0;
  for (int u = 0; u < d->st_n; u++) {
    if (n_new[u] < 0) { set_error("stream %d: negative frame count", u); return JB200_ERR_ARG; }
    if (d->st_done[u] && n_new[u] > 0) { set_error("stream %d has ended; restart it before feeding more frames", u); return JB200_ERR_ARG; }
    if (d->st_t[u] + n_new[u] > d->st_cap) { set_error("stream %d: %d frames exceed the per-stream capacity %d", u, d->st_t[u] + n_new[u], d->st_cap); return JB200_ERR_CAPACITY; }
    tot += n_new[u];
  }
  if (tot > d->max_frames) { set_error("%d new frames exceed decoder capacity %d", tot, d->max_frames); return JB200_ERR_CAPACITY; }
  *total = tot;
  return JB200_OK;
}

extern "C" int jb200_stream_feed_host(jb200_decoder *d, const float *feats, const int32_t *n_new, const uint8_t *last, int want_interim) {
  int tot = 0;
  int rc = stream_check(d, n_new, &tot); if (rc) return rc;
  if (tot > 0 && !feats) { set_error("null feats"); return JB200_ERR_ARG; }
  JB_CUDA(cudaSetDevice(d->device));
  if (tot > 0) {
    JB_CUDA(cudaMemcpyAsync(d->d_feats, feats, sizeof(float) * (size_t)tot * d->dim, cudaMemcpyHostToDevice, d->stream));
    rc = d->dnn ? dnn_forward_device(d->dnn, d->d_feats, tot, d->d_rows, d->row_stride, d->stream)
                : gmm_launch_states(d->am, d->d_feats, tot, d->d_rows, d->row_stride, d->stream, nullptr, nullptr, 0);
    if (rc) return rc;
  }
  return stream_advance(d, n_new, last, want_interim);
}

extern "C" int jb200_stream_feed_scores_host(jb200_decoder *d, const float *scores, const int32_t *n_new, const uint8_t *last, int want_interim) {
  int tot = 0;
  int rc = stream_check(d, n_new, &tot); if (rc) return rc;
  if (tot > 0 && !scores) { set_error("null scores"); return JB200_ERR_ARG; }
  JB_CUDA(cudaSetDevice(d->device));
  if (tot > 0)
    JB_CUDA(cudaMemcpy2DAsync(d->d_rows, sizeof(float) * d->row_stride, scores, sizeof(float) * d->S, sizeof(float) * d->S, tot, cudaMemcpyHostToDevice, d->stream));
  return stream_advance(d, n_new, last, want_interim);
}

extern "C" int jb200_stream_restart(jb200_decoder *d, int stream) {
  if (!d || !d->stream_mode || stream < 0 || stream >= d->st_n) { set_error("jb200_stream_restart: bad argument"); return JB200_ERR_ARG; }
  JB_CUDA(cudaSetDevice(d->device));
  if (d->st_started[stream] && !d->st_done[stream]) {
    const size_t tot = (size_t)d->P.n_nodes;
    fill_slots_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, d->stream>>>(d->P.slots + (size_t)stream * d->P.n_nodes, tot);
    JB_LAUNCH_CHECK();
    JB_CUDA(cudaStreamSynchronize(d->stream));
  }
  d->st_t[stream] = 0; d->st_started[stream] = 0; d->st_done[stream] = 0;
  return JB200_OK;
}

extern "C" int jb200_stream_status(jb200_decoder *d, int stream, int32_t *frames_done, int32_t *alive, int32_t *ended) {
  if (!d || !d->stream_mode || stream < 0 || stream >= d->st_n) { set_error("jb200_stream_status: bad argument"); return JB200_ERR_ARG; }
  if (frames_done) *frames_done = d->st_t[stream];
  if (alive) *alive = (!d->st_started[stream] || d->st_done[stream]) ? 1 : (d->h_state[stream].stopped < 0);
  if (ended) *ended = d->st_done[stream];
  return JB200_OK;
}

extern "C" int jb200_stream_partial(jb200_decoder *d, int stream, int32_t *words, int max_words, int32_t *n_words, float *score, int32_t *frame) {
  if (!d || !d->stream_mode || stream < 0 || stream >= d->st_n || !n_words) { set_error("jb200_stream_partial: bad argument"); return JB200_ERR_ARG; }
  if (!d->st_started[stream] || d->st_done[stream]) { *n_words = 0; if (score) *score = JB200_LOG_ZERO; if (frame) *frame = -1; return JB200_OK; }
  const UttState &st = d->h_state[stream];
  const int n = std::min(st.interim_nwords, std::max(max_words, 0));
  for (int i = 0; i < n && words; i++) words[i] = d->h_interim_words[(size_t)stream * MAX_WORDS + i];
  *n_words = words ? n : st.interim_nwords;
  if (score) *score = st.interim_score;
  if (frame) *frame = st.interim_frame;
  return JB200_OK;
}

extern "C" int jb200_stream_result(jb200_decoder *d, int stream, const jb200_utt_result **utt, const jb200_atom **atoms, const int32_t **words) {
  if (!d || !d->stream_mode || stream < 0 || stream >= d->st_n) { set_error("jb200_stream_result: bad argument"); return JB200_ERR_ARG; }
  if (!d->st_done[stream]) { set_error("stream %d has not ended", stream); return JB200_ERR_ARG; }
  if (utt) *utt = d->h_results + stream;
  if (atoms) *atoms = d->h_atoms;        // index with utt->atom_offset, as for a batch
  if (words) *words = d->h_words;        // index with utt->word_offset
  return JB200_OK;
}

extern "C" int jb200_decoder_pipeline_info(jb200_decoder *d, int32_t *n_slices, float *score_busy_ms) {
  if (!d) { set_error("null decoder"); return JB200_ERR_ARG; }
  if (n_slices) *n_slices = d->last_piped ? d->n_chunks : 1;
  if (score_busy_ms) *score_busy_ms = d->last_piped ? d->last_score_busy_ms : 0.0f;
  return JB200_OK;
}

extern "C" int jb200_decoder_set_pipeline(jb200_decoder *d, int frames_per_slice) {
  if (!d || frames_per_slice < 0) { set_error("jb200_decoder_set_pipeline: bad argument"); return JB200_ERR_ARG; }
  d->pipe_frames = frames_per_slice;
  return JB200_OK;
}

extern "C" int jb200_decoder_results(jb200_decoder *d, const jb200_utt_result **utts, const jb200_atom **atoms, const int32_t **words) {
  if (!d) { set_error("null decoder"); return JB200_ERR_ARG; }
  if (!d->fetched) { int rc = jb200_decoder_fetch(d); if (rc) return rc; }
  if (utts) *utts = d->h_results;
  if (atoms) *atoms = d->h_atoms;
  if (words) *words = d->h_words;
  return JB200_OK;
}

extern "C" int jb200_decoder_last_timing(jb200_decoder *d, float ms[4]) {
  if (!d || !ms) { set_error("bad argument"); return JB200_ERR_ARG; }
  for (int i = 0; i < 4; i++) ms[i] = d->last_ms[i];
  return JB200_OK;
}

extern "C" int jb200_decoder_frame_counts(jb200_decoder *d, int u, int32_t *counts, int max_frames) {
  if (!d || !counts || u < 0 || u >= d->last_n) { set_error("bad argument"); return JB200_ERR_ARG; }
  JB_CUDA(cudaSetDevice(d->device));
  const int f0 = d->h_frame_off[u], T = d->h_frame_off[u + 1] - f0;
  const int n = T < max_frames ? T : max_frames;
  JB_CUDA(cudaMemcpy(counts, d->P.counts + (size_t)f0 * 2, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost));
  return JB200_OK;
}
