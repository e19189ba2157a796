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
      for (int k = 0; k < extract; k++) { checks++; if (outv[k].id != R[n - k].id) { mism++; break; } }
      for (int k = 0; k < extract; k++) K2[n - k] = outv2[k];
      for (int i = 1; i <= n; i++) { checks++; if (K2[i].id != R[i].id) { mism++; break; } }
    } else {
      reference_select<false>(R, n, extract);
      kernel_select<false>(K, n, extract, maxt, NEG, outv);
      for (int i = 1; i <= need; i++) { checks++; if (K[i].id != R[i].id) { mism++; break; } }
      for (int k = 0; k < extract; k++) { checks++; if (outv[k].id != R[n - k].id) { mism++; break; } }
    }
  }
  printf("trials %d, element checks %ld, mismatches %ld\n", trials, checks, mism);
  return mism != 0;
}
