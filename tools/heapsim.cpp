// tools/heapsim.cpp -- CPU model of the beam cut as beam.cu runs it (heap_pad_sentinels + heap_extract_fast with the
// loser cut), checked against the reference's loop (sort_token_upward / _downward, libjulius/src/beam.c:1342-1447).
//
//   g++ -O2 -o /tmp/heapsim tools/heapsim.cpp && /tmp/heapsim [score divisor, small = many exact ties] [trials]
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

int main(int argc, char **argv) {
  const double divisor = argc > 1 ? atof(argv[1]) : 7.0;
  const int trials = argc > 2 ? atoi(argv[2]) : 3000;
  srand(1);
  long mism = 0, checks = 0;
  for (int tr = 0; tr < trials; tr++) {
    const int n = 3 + rand() % 2600, need = 1 + rand() % (n - 1);
    const int maxt = ((std::max(n, 64) + 63 + rand() % 500) + 3) & ~3;
    std::vector<Ent> base(2 * maxt + 8, Ent{0, 0.0f});
    std::vector<float> scores;
    for (int i = 1; i <= n; i++) { base[i] = Ent{i - 1, -(float)(rand() % 20000) / (float)divisor}; scores.push_back(base[i].v); }
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
  printf("trials %d, element checks %ld, mismatches %ld\n", trials, checks, mism);
  return mism != 0;
}
