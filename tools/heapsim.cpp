// tools/heapsim.cpp -- CPU model of the beam cut as beam.cu runs it (heap_pad_sentinels + heap_extract_fast with the
// loser cut), checked against the reference's loop (sort_token_upward / _downward, libjulius/src/beam.c:1342-1447).
//
//   g++ -O2 -o /tmp/heapsim tools/heapsim.cpp && /tmp/heapsim [score divisor] [trials] [distinct scores, small = many exact ties]
//
// What is compared, per trial (n tokens, `need` survivors, random scores on a coarse grid so that ties are routine):
//   upward selects   (need <  n-need): the order in which the `need` largest elements are extracted;
//   downward selects (need >= n-need): the arrangement of the `need` elements left in the heap;
//   and, without the cut, the complete final array (what the multipath kernel's select #1 needs).
// The model mirrors the kernel statement by statement: freed tail slots and everything up to the last child slot hold
// a sentinel, so the loop has no bounds tests; the address of the next children pair is clamped to a sentinel pair;
// max-heap sifts stop when the larger child is below `lose_below` (a lower bound of the need-th largest score).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>

struct Ent { int id; float v; };
static const float NEG = -INFINITY, POS = INFINITY;

template <bool MAXHEAP> static bool hcmp(float a, float b) { return MAXHEAP ? (a < b) : (a > b); }      // "child < child+1"
template <bool MAXHEAP> static bool hstop(float s, float c) { return MAXHEAP ? (s >= c) : (s <= c); }   // "STVAL >= SVAL(child)"

template <bool MAXHEAP>
static void sift_down(std::vector<Ent> &A, int start, int n) {          // the reference's inner loop, bounds tests and all
  Ent s = A[start];
  int parent = start, child;
  while ((child = parent * 2) <= n) {
    if (child < n && hcmp<MAXHEAP>(A[child].v, A[child + 1].v)) child++;
    if (hstop<MAXHEAP>(s.v, A[child].v)) break;
    A[parent] = A[child];
    parent = child;
  }
  A[parent] = s;
}

// reference: build + extract in place; the k-th extracted root ends in slot n-k
template <bool MAXHEAP>
static void reference_select(std::vector<Ent> &A, int n, int extract) {
  for (int root = n / 2; root >= 1; root--) sift_down<MAXHEAP>(A, root, n);
  int m = n;
  while (m > n - extract) {
    Ent s = A[m];
    A[m] = A[1];
    m--;
    if (m < 1) break;
    A[1] = s;
    sift_down<MAXHEAP>(A, 1, m);
  }
}

// the kernel's formulation; outv[k] = k-th extracted root
template <bool MAXHEAP>
static void kernel_select(std::vector<Ent> &A, int n, int extract, int maxt, float lose_below, std::vector<Ent> &outv) {
  const float sent = MAXHEAP ? NEG : POS;
  for (int i = n + 1; i <= std::min(2 * n + 1, maxt + 1); i++) A[i] = Ent{0, sent};      // heap_pad_sentinels
  A[maxt + 2] = A[maxt + 3] = Ent{0, sent};
  for (int root = n / 2; root >= 1; root--) sift_down<MAXHEAP>(A, root, n);              // heap_build (level-parallel on the GPU)
  const int cap = maxt / 2 + 1;                                                          // pair index of (maxt+2, maxt+3)
  int mslot = n;
  outv.assign(extract, Ent{-1, 0});
  for (int x = 0; x < extract; x++) {
    const Ent s = A[mslot];
    A[mslot] = Ent{0, sent};                       // the slot leaves the heap
    outv[x] = A[1];
    mslot--;
    int slot = 1, cur = 1;                         // parent slot, pair index of its children (slots 2*cur, 2*cur+1)
    Ent x0 = A[2], y0 = A[3];
    while (true) {
      const bool right = hcmp<MAXHEAP>(x0.v, y0.v);
      const int child = 2 * cur + (right ? 1 : 0);
      const int ncur = std::min(child, cap);       // speculative load address (clamped)
      const Ent nx = A[2 * ncur], ny = A[2 * ncur + 1];
      const Ent c = right ? y0 : x0;
      if (hstop<MAXHEAP>(s.v, c.v) || (MAXHEAP && c.v < lose_below)) break;
      A[slot] = c;
      slot = child;
      cur = ncur;
      x0 = nx; y0 = ny;
    }
    A[slot] = s;
  }
}


// ---- the pipelined formulation (heap_extract_pipe in beam.cu) -------------------------------------------------------
// One warp, lock-step "ticks".  Extraction x is owned by lane x mod NL; an extraction in flight moves its hole down
// exactly one tree level per tick; a new extraction starts at the earliest two ticks after the previous one (so it reads
// level L+1 one tick after its predecessor wrote it) and only when no extraction in flight has its hole on an ancestor
// of (or at) the tail slot it is about to take -- such an extraction could still place its own s into that slot, and it
// still compares against the slot's real content.  The value an extraction writes into the root is the next one's
// output.  State is updated in the order the kernel uses inside a tick: (1) all lanes in flight read their children
// pair, (2) they write their hole and move, (3) the start decision is taken on the holes as they are now; the starting
// lane reads its s and retires the tail slot, and works on the root from the next tick on.
template <bool MAXHEAP>
static long pipelined_select(std::vector<Ent> &A, int n, int extract, int maxt, float lose_below, std::vector<Ent> &outv,
                             int NL = 16) {
  const float sent = MAXHEAP ? NEG : POS;
  for (int i = n + 1; i <= std::min(2 * n + 1, maxt + 1); i++) A[i] = Ent{0, sent};
  A[maxt + 2] = A[maxt + 3] = Ent{0, sent};
  for (int root = n / 2; root >= 1; root--) sift_down<MAXHEAP>(A, root, n);
  const int cap = maxt / 2 + 1;
  struct Lane { bool act = false; int x = 0, slot = 0, cur = 0; Ent s{0, 0}; };
  std::vector<Lane> L(NL);
  outv.assign(extract + 1, Ent{-1, 0});
  if (extract > 0) outv[0] = A[1];
  int next_x = 0, wait = 0; long ticks = 0;
  while (true) {
    bool any = false;
    for (auto &l : L) any |= l.act;
    if (next_x >= extract && !any) break;
    ticks++;
    // (1) every extraction in flight reads the children pair of its hole
    std::vector<Ent> rx(NL), ry(NL);
    for (int k = 0; k < NL; k++) if (L[k].act) { rx[k] = A[2 * L[k].cur]; ry[k] = A[2 * L[k].cur + 1]; }
    // (2) every extraction in flight fills its hole and moves one level down (or ends)
    for (int k = 0; k < NL; k++) if (L[k].act) {
      Lane &l = L[k];
      const bool right = hcmp<MAXHEAP>(rx[k].v, ry[k].v);
      const Ent c = right ? ry[k] : rx[k];
      const bool stop = hstop<MAXHEAP>(l.s.v, c.v) || (MAXHEAP && c.v < lose_below);
      const Ent put = stop ? l.s : c;
      A[l.slot] = put;
      if (l.slot == 1) outv[l.x + 1] = put;            // what goes into the root is the next extraction's output
      if (stop) l.act = false;
      else { const int child = 2 * l.cur + (right ? 1 : 0); l.slot = child; l.cur = std::min(child, cap); }
    }
    // (3) start decision on the holes as they are after this tick's move; the new extraction's first level is the
    //     next tick's business
    int started = -1;
    if (--wait <= 0 && next_x < extract) {
      const int ms = n - next_x, ln = next_x % NL;
      bool blocked = L[ln].act;
      // a tail slot holding a loser is never a hole and never decides anything (loser cut), and a loser stays a loser
      const bool loser = MAXHEAP && A[ms].v < lose_below;
      for (int k = 0; k < NL && !blocked && !loser; k++) if (L[k].act) {
        int a = ms;
        while (a > L[k].slot) a >>= 1;
        if (a == L[k].slot) blocked = true;
      }
      if (!blocked) {
        L[ln].s = A[ms]; A[ms] = Ent{0, sent}; L[ln].x = next_x;
        started = ln; next_x++; wait = 2;
      }
    }
    if (started >= 0) { L[started].act = true; L[started].slot = 1; L[started].cur = 1; }
  }
  return ticks;
}

// ---- the closed form of an upward select (heap_select_closed in beam.cu) ------------------------------------------------
// When every sift of the extraction loop ends on a loser (an element that is never extracted), an extraction is a pure
// "pull-up": the hole at the root is filled by the larger child (the left one on a tie), and so on down.  Two elements of
// equal score then keep their relative PRE-ORDER position in the tree for ever (the one in the right subtree of their
// lowest common ancestor could only overtake by being strictly larger than everything in the left subtree), the root is
// first in pre-order, hence:   extraction order = (score descending, pre-order position in the built heap ascending).
// A re-inserted WINNER (the tail slot taken by an extraction holds one of the elements that will be extracted) sinks from
// the root instead and may end up ahead of elements it ties with; it cannot disturb the order of anybody else.  Such an
// element e sits in a tail slot p of the built heap (tail slots are leaves; nothing is ever promoted into a leaf, so what
// a tail slot holds when it is taken is its original content or an earlier extraction's s, itself a tail content), and
// when slot p is taken at step k = n-p+1 the k-1 elements extracted so far and the d = depth(p) elements on the slots
// above p are all ahead of e in the order above.  So if   rank(e) < k + d   for every tail element e that ties with
// another candidate, no re-insertion can matter and the closed form is exact; otherwise the caller replays the loop.
struct CfEnt { float v; int pre; int id; int slot; };
static long g_fail_bound = 0, g_fail_lander = 0, g_gap_sum = 0;
static int cf_subtree_size(int c, int n) {
  int H = 31 - __builtin_clz(n), dc = 31 - __builtin_clz(c);
  if (dc > H) return 0;
  const int full = (1 << (H - dc)) - 1;                       // levels dc .. H-1
  const long first = (long)c << (H - dc), width = 1L << (H - dc);
  long last_cnt = (long)n - first + 1; if (last_cnt < 0) last_cnt = 0; if (last_cnt > width) last_cnt = width;
  return full + (int)last_cnt;
}
static int cf_preorder(int h, int n) {
  int pre = 0, cur = 1;
  const int d = 31 - __builtin_clz(h);
  for (int b = d - 1; b >= 0; b--) {
    const int bit = (h >> b) & 1;
    pre += 1;
    if (bit) pre += cf_subtree_size(cur * 2, n);
    cur = cur * 2 + bit;
  }
  return pre;
}
// returns true when the closed form applies; order[k] = id of the k-th extracted
static bool closed_form_select(const std::vector<Ent> &H0, int n, int need, float lose_below, std::vector<int> &order, int level = 3) {
  std::vector<CfEnt> c;
  for (int h = 1; h <= n; h++) if (H0[h].v >= lose_below) c.push_back(CfEnt{H0[h].v, cf_preorder(h, n), H0[h].id, h});
  if ((int)c.size() < need) return false;
  std::sort(c.begin(), c.end(), [](const CfEnt &a, const CfEnt &b) { return a.v != b.v ? a.v > b.v : a.pre < b.pre; });
  const float theta = c[need - 1].v;
  // rank (1-based) of the element in each heap slot, 0 = not a candidate with score >= theta
  std::vector<int> rank_of_slot(n + 2, 0);
  size_t nw = 0;
  for (size_t i = 0; i < c.size() && c[i].v >= theta; i++) { rank_of_slot[c[i].slot] = (int)i + 1; nw = i + 1; }
  // PR: tail elements that may still be in their slot when it is taken (the necessary condition rank >= k + d)
  struct Pr { int k; float v; int slot; };
  std::vector<Pr> pr;
  for (size_t i = 0; i < nw; i++) {
    if (c[i].slot < n - need + 1) continue;
    const int k = n - c[i].slot + 1, d = 31 - __builtin_clz(c[i].slot);
    if ((int)i + 1 >= k + d) pr.push_back(Pr{k, c[i].v, c[i].slot});
  }
  for (size_t i = 0; i < nw; i++) {
    const bool tied = (i > 0 && c[i - 1].v == c[i].v) || (i + 1 < c.size() && c[i + 1].v == c[i].v);
    if (!tied || c[i].slot < n - need + 1) continue;
    const int p = c[i].slot, k = n - p + 1, d = 31 - __builtin_clz(p);
    if ((int)i + 1 < k + d) continue;                      // cannot be in its slot any more
    if (level < 3) return false;
    // the parent's element f leaves slot q = p/2 by step rank(f) - (d-1); the hole it leaves takes e unless the sibling
    // leaf holds something ahead of e (then e follows when that one leaves q), or a re-inserted element lands in between
    const int q = p >> 1, sib = p ^ 1;
    const int rf = rank_of_slot[q];
    if (rf == 0) return false;                             // (cannot happen: the parent of a winner is a winner)
    int bound = rf - d + 1;
    if (sib <= n && rank_of_slot[sib] != 0) {
      const float gv = H0[sib].v;
      const bool g_beats = (gv > c[i].v) || (gv == c[i].v && sib < p);
      if (g_beats) bound = std::max(bound, rank_of_slot[sib] - d + 1);
    }
    if (!(bound < k)) { g_fail_bound++; g_gap_sum += bound - k; return false; }
    const float fv = H0[q].v;
    for (const Pr &t : pr) if (t.slot != p && t.k < k && t.v >= c[i].v && t.v <= fv) { g_fail_lander++; return false; }
  }
  order.resize(need);
  for (int k = 0; k < need; k++) order[k] = c[k].id;
  return true;
}

// heapsim --dump <file>: the beam cuts of a real decode (tools/dump_heaps.py): how often the closed form applies, and
// that it is right whenever it does
static int dump_mode(const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); return 2; }
  long nup = 0, used2 = 0, used3 = 0, wrong = 0, right_anyway = 0, n_sus_frames = 0, n_taken = 0, sum_leave_step = 0;
  int hdr[2];
  while (fread(hdr, 4, 2, f) == 2) {
    const int n = hdr[0], need = hdr[1];
    std::vector<Ent> A(n + 2, Ent{0, 0.0f});
    for (int i = 1; i <= n; i++) { A[i].id = i - 1; if (fread(&A[i].v, 4, 1, f) != 1) return 2; }
    if (!(need < n - need)) continue;
    nup++;
    std::vector<Ent> R = A, H0 = A;
    reference_select<true>(R, n, need);
    for (int root = n / 2; root >= 1; root--) sift_down<true>(H0, root, n);
    std::vector<float> sc; for (int i = 1; i <= n; i++) sc.push_back(A[i].v);
    std::sort(sc.begin(), sc.end(), std::greater<float>());
    const float lose_below = sc[need - 1] - 0.25f;
    std::vector<int> o2, o3;
    if (!closed_form_select(H0, n, need, lose_below, o2, 2)) {
      // suspects: tied tail candidates; replay the plain loop and note the step after which none of them sits in its slot
      std::vector<std::pair<int,int>> sus;   // (slot, id)
      { std::vector<float> ws; for (int i = 1; i <= n; i++) if (H0[i].v >= sc[need - 1]) ws.push_back(H0[i].v);
        std::sort(ws.begin(), ws.end());
        for (int p = n - need + 1; p <= n; p++) if (H0[p].v >= sc[need - 1]) {
          auto r = std::equal_range(ws.begin(), ws.end(), H0[p].v);
          if (r.second - r.first > 1) sus.push_back({p, H0[p].id});
        } }
      std::vector<Ent> W = H0; int m = n, step = 0, taken = 0, last_leave = 0;
      std::vector<char> gone(sus.size(), 0);
      while (m > n - need) {
        step++;
        for (size_t q = 0; q < sus.size(); q++) if (!gone[q] && sus[q].first == m && W[m].id == sus[q].second) { taken = 1; }
        Ent s0 = W[m]; W[m] = W[1]; m--; if (m < 1) break; W[1] = s0; sift_down<true>(W, 1, m);
        for (size_t q = 0; q < sus.size(); q++) if (!gone[q] && (sus[q].first > m || W[sus[q].first].id != sus[q].second)) { gone[q] = 1; last_leave = step; }
        bool all = true; for (char g : gone) all = all && g;
        if (all) break;
      }
      n_sus_frames++; if (taken) n_taken++; else sum_leave_step += last_leave;
    }
    if (closed_form_select(H0, n, need, lose_below, o2, 2)) used2++;
    if (closed_form_select(H0, n, need, lose_below, o3, 3)) {
      used3++;
      for (int k = 0; k < need; k++) if (o3[k] != R[n - k].id) { wrong++; break; }
    }
  }
  printf("upward selects %ld: closed form applies in %ld (%.1f%%) with the rank test alone, %ld (%.1f%%) with the parent/sibling bound; wrong %ld\n",
         nup, used2, 100.0 * used2 / nup, used3, 100.0 * used3 / nup, wrong);
  printf("failures: bound %ld (mean bound-k %.1f), lander %ld\n", g_fail_bound, g_fail_bound ? (double)g_gap_sum / g_fail_bound : 0.0, g_fail_lander);
  printf("frames with suspects %ld: a suspect was taken from its slot in %ld; otherwise all suspects had left their leaves after %.1f extractions on average\n", n_sus_frames, n_taken, n_sus_frames > n_taken ? (double)sum_leave_step / (n_sus_frames - n_taken) : 0.0);
  (void)right_anyway;
  return wrong != 0;
}

int main(int argc, char **argv) {
  if (argc > 2 && std::string(argv[1]) == "--dump") return dump_mode(argv[2]);
  const double divisor = argc > 1 ? atof(argv[1]) : 7.0;
  const int trials = argc > 2 ? atoi(argv[2]) : 3000;
  const int modulus = argc > 3 ? atoi(argv[3]) : 20000;        // number of distinct scores: small = many exact ties
  srand(1);
  long mism = 0, checks = 0, cf_total = 0, cf_used = 0, cf_bad = 0;
  for (int tr = 0; tr < trials; tr++) {
    const int n = 3 + rand() % 2600, need = 1 + rand() % (n - 1);
    const int maxt = ((std::max(n, 64) + 63 + rand() % 500) + 3) & ~3;
    std::vector<Ent> base(2 * maxt + 8, Ent{0, 0.0f});
    std::vector<float> scores;
    for (int i = 1; i <= n; i++) { base[i] = Ent{i - 1, -(float)(rand() % modulus) / (float)divisor}; scores.push_back(base[i].v); }
    const bool upward = need < n - need;
    const int extract = upward ? need : n - need;
    std::vector<Ent> R = base, K = base, K2 = base, outv, outv2;
    if (upward) {
      std::sort(scores.begin(), scores.end(), std::greater<float>());
      // any lower bound of the need-th largest score is legal; use one a little below it, as the histogram does
      const float lose_below = scores[need - 1] - (float)(rand() % 3) * 0.5f;
      reference_select<true>(R, n, extract);
      kernel_select<true>(K, n, extract, maxt, lose_below, outv);
      kernel_select<true>(K2, n, extract, maxt, NEG, outv2);           // no cut: the whole array must agree
      { std::vector<Ent> H0 = base; for (int root = n / 2; root >= 1; root--) sift_down<true>(H0, root, n);
        std::vector<int> ord;
        cf_total++;
        if (closed_form_select(H0, n, need, lose_below, ord)) {
          cf_used++;
          for (int k = 0; k < extract; k++) { checks++; if (ord[k] != R[n - k].id) { mism++; cf_bad++; break; } }
        } }
      { std::vector<Ent> P = base, P2 = base, po, po2;
        pipelined_select<true>(P, n, extract, maxt, lose_below, po);
        for (int k = 0; k < extract; k++) { checks++; if (po[k].id != R[n - k].id) { mism++; break; } }
        pipelined_select<true>(P2, n, extract, maxt, NEG, po2, 8);
        for (int k = 0; k < extract; k++) P2[n - k] = po2[k];
        for (int i = 1; i <= n; i++) { checks++; if (P2[i].id != R[i].id) { mism++; break; } } }
      for (int k = 0; k < extract; k++) { checks++; if (outv[k].id != R[n - k].id) { mism++; break; } }
      for (int k = 0; k < extract; k++) K2[n - k] = outv2[k];
      for (int i = 1; i <= n; i++) { checks++; if (K2[i].id != R[i].id) { mism++; break; } }
    } else {
      reference_select<false>(R, n, extract);
      kernel_select<false>(K, n, extract, maxt, NEG, outv);
      { std::vector<Ent> P = base, po;
        pipelined_select<false>(P, n, extract, maxt, NEG, po);
        for (int i = 1; i <= need; i++) { checks++; if (P[i].id != R[i].id) { mism++; break; } }
        for (int k = 0; k < extract; k++) { checks++; if (po[k].id != R[n - k].id) { mism++; break; } } }
      for (int i = 1; i <= need; i++) { checks++; if (K[i].id != R[i].id) { mism++; break; } }
      for (int k = 0; k < extract; k++) { checks++; if (outv[k].id != R[n - k].id) { mism++; break; } }
    }
  }
  printf("trials %d, element checks %ld, mismatches %ld; closed form used in %ld of %ld upward selects, wrong %ld\n", trials, checks, mism, cf_used, cf_total, cf_bad);
  return mism != 0;
}
