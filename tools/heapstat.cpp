// tools/heapstat.cpp -- statistics of the beam cuts of a real decode (input: tools/dump_heaps.py), and a CPU check of the
// closed form for the extraction order of sort_token_upward (libjulius/src/beam.c:1342-1386) against the plain loop.
//   g++ -O2 -o /tmp/heapstat tools/heapstat.cpp && /tmp/heapstat /tmp/hd/tri20k.heaps
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <map>

struct Ent { int id; float v; };

template <bool MAXHEAP> static bool hcmp(float a, float b) { return MAXHEAP ? (a < b) : (a > b); }
template <bool MAXHEAP> static bool hstop(float s, float c) { return MAXHEAP ? (s >= c) : (s <= c); }
template <bool MAXHEAP>
static int sift_down(std::vector<Ent> &A, int start, int n) {
  Ent s = A[start];
  int parent = start, child, lv = 0;
  while ((child = parent * 2) <= n) {
    if (child < n && hcmp<MAXHEAP>(A[child].v, A[child + 1].v)) child++;
    lv++;
    if (hstop<MAXHEAP>(s.v, A[child].v)) break;
    A[parent] = A[child];
    parent = child;
  }
  A[parent] = s;
  return lv;
}
template <bool MAXHEAP>
static void build(std::vector<Ent> &A, int n) { for (int r = n / 2; r >= 1; r--) sift_down<MAXHEAP>(A, r, n); }

// pre-order rank of heap slot p in a complete binary tree with n slots
static int subtree_size(int p, int n) {
  // number of slots in the subtree of p among 1..n
  int cnt = 0; long lo = p, hi = p;
  while (lo <= n) { cnt += (int)(std::min<long>(hi, n) - lo + 1); lo = lo * 2; hi = hi * 2 + 1; }
  return cnt;
}
static int preorder(int p, int n) {
  // rank = number of slots visited before p in a root-left-right walk
  int rank = 0;
  // walk from the root to p
  int depth = 31 - __builtin_clz(p);
  int cur = 1;
  for (int d = depth - 1; d >= 0; d--) {
    const int bit = (p >> d) & 1;
    rank += 1;                                   // cur itself
    if (bit) rank += subtree_size(cur * 2, n);   // whole left subtree
    cur = cur * 2 + bit;
  }
  return rank;
}

// ticks of the pipelined replay (tools/heapsim.cpp: pipelined_select) on a built max-heap, loser cut at lose_below
static long pipe_ticks(std::vector<Ent> A, int n, int extract, float lose_below, long *stalls) {
  const float NEG = -INFINITY;
  A.resize(2 * n + 8, Ent{0, NEG});
  for (int i = n + 1; i < (int)A.size(); i++) A[i] = Ent{0, NEG};
  const int NL = 16;
  struct Lane { bool act = false; int slot = 0, cur = 0; Ent s{0, 0}; };
  std::vector<Lane> L(NL);
  int next_x = 0, wait = 0; long ticks = 0;
  while (true) {
    bool any = false; for (auto &l : L) any |= l.act;
    if (next_x >= extract && !any) break;
    ticks++;
    std::vector<Ent> rx(NL), ry(NL);
    for (int k = 0; k < NL; k++) if (L[k].act) { rx[k] = A[2 * L[k].cur]; ry[k] = A[2 * L[k].cur + 1]; }
    int started = -1;
    if (--wait <= 0 && next_x < extract) {
      const int ms = n - next_x, ln = next_x % NL;
      bool blocked = L[ln].act;
      const bool loser = A[ms].v < lose_below;
      for (int k = 0; k < NL && !blocked && !loser; k++) if (L[k].act) { int a = ms; while (a > L[k].slot) a >>= 1; if (a == L[k].slot) blocked = true; }
      if (!blocked) { L[ln].s = A[ms]; A[ms] = Ent{0, NEG}; started = ln; next_x++; wait = 2; } else (*stalls)++;
    }
    for (int k = 0; k < NL; k++) if (L[k].act) {
      Lane &l = L[k];
      const bool right = rx[k].v < ry[k].v;
      const Ent c = right ? ry[k] : rx[k];
      const bool stop = (l.s.v >= c.v) || (c.v < lose_below);
      A[l.slot] = stop ? l.s : c;
      if (stop) l.act = false; else { const int child = 2 * l.cur + (right ? 1 : 0); l.slot = child; l.cur = child; }
    }
    if (started >= 0) { L[started].act = true; L[started].slot = 1; L[started].cur = 1; }
  }
  return ticks;
}

int main(int argc, char **argv) {
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 1;
  long nsel = 0, nup = 0, ndown = 0, sum_n = 0, sum_lv = 0, sum_x = 0;
  long c1_ok = 0, cf_ok = 0, cf_ok_given_c1 = 0, c1_total = 0, c2_ok = 0, cf_ok_given_c2 = 0;
  long sum_ticks = 0, sum_stalls = 0;
  long sum_tail_win = 0, sum_tie_pairs = 0, sum_reins = 0, sum_reins_tied = 0;
  std::map<int, long> badhist;
  int hdr[2];
  while (fread(hdr, 4, 2, f) == 2) {
    const int n = hdr[0], need = hdr[1];
    std::vector<Ent> A(n + 2);
    for (int i = 1; i <= n; i++) { A[i].id = i - 1; if (fread(&A[i].v, 4, 1, f) != 1) return 2; }
    nsel++; sum_n += n;
    const bool upward = need < n - need;
    if (!upward) { ndown++; continue; }
    nup++;
    // reference
    std::vector<Ent> R = A;
    build<true>(R, n);
    std::vector<Ent> H0 = R;
    std::vector<int> ref_order;
    int m = n; long reins_w = 0;
    std::vector<float> sorted;
    for (int i = 1; i <= n; i++) sorted.push_back(A[i].v);
    std::sort(sorted.begin(), sorted.end(), std::greater<float>());
    const float theta = sorted[need - 1];
    sum_ticks += pipe_ticks(H0, n, need, theta - 0.25f, &sum_stalls);
    std::vector<int> reinserted;     // ids of winners re-inserted
    while (m > n - need) {
      Ent s = R[m]; ref_order.push_back(R[1].id); R[m] = R[1]; m--;
      if (m < 1) break;
      if (s.v >= theta) { reins_w++; reinserted.push_back(s.id); }
      R[1] = s; sum_lv += sift_down<true>(R, 1, m); sum_x++;
    }
    sum_reins += reins_w;
    // closed form: sort winners by (score desc, preorder(H0 slot) asc)
    struct W { float v; int pre; int id; };
    std::vector<W> w;
    for (int p = 1; p <= n; p++) if (H0[p].v >= theta) w.push_back(W{H0[p].v, preorder(p, n), H0[p].id});
    std::sort(w.begin(), w.end(), [](const W &a, const W &b) { return a.v != b.v ? a.v > b.v : a.pre < b.pre; });
    bool same = true;
    for (int k = 0; k < need; k++) if (w[k].id != ref_order[k]) { same = false; break; }
    // condition C1': no winner in the original tail slots ties with another winner
    std::map<float, int> mult;
    for (auto &x : w) mult[x.v]++;
    long tie_pairs = 0; for (auto &kv : mult) tie_pairs += (long)kv.second * (kv.second - 1) / 2;
    sum_tie_pairs += tie_pairs;
    bool c1 = true; long tail_w = 0, bad = 0;
    for (int p = n - need + 1; p <= n; p++) if (H0[p].v >= theta) { tail_w++; if (mult[H0[p].v] > 1) { c1 = false; bad++; } }
    sum_tail_win += tail_w;
    badhist[(int)std::min<long>(bad, 10)]++;
    c1_total++;
    // condition C2: a tied tail winner e in slot p (taken at step k = n-p+1) can only be re-inserted if its d = depth(p)
    // ancestors and the k-1 elements extracted before are all ahead of it:  rank(e) >= k + d
    {
      std::map<int, int> rank_of;   // id -> 1-based rank in the closed-form order
      for (size_t i = 0; i < w.size(); i++) rank_of[w[i].id] = (int)i + 1;
      bool c2 = true;
      for (int p = n - need + 1; p <= n; p++) if (H0[p].v >= theta && mult[H0[p].v] > 1) {
        const int k = n - p + 1, d = 31 - __builtin_clz(p);
        if (rank_of[H0[p].id] >= k + d) c2 = false;
      }
      if (c2) { c2_ok++; if (same) cf_ok_given_c2++; }
    }
    if (c1) { c1_ok++; if (same) cf_ok_given_c1++; }
    if (same) cf_ok++;
  }
  printf("selects %ld (upward %ld, downward %ld), mean n %.1f\n", nsel, nup, ndown, (double)sum_n / nsel);
  printf("upward: levels/extraction %.2f, extractions/select %.1f\n", (double)sum_lv / sum_x, (double)sum_x / nup);
  printf("pipelined replay: %.0f ticks/select, %.1f stalled ticks/select\n", (double)sum_ticks / nup, (double)sum_stalls / nup);
  printf("winners in tail slots of H0 per select %.1f, re-inserted winners per select %.1f, tied winner pairs per select %.1f\n",
         (double)sum_tail_win / nup, (double)sum_reins / nup, (double)sum_tie_pairs / nup);
  printf("C1' holds in %ld of %ld (%.1f%%); closed form correct in %ld (%.1f%%); correct given C1' %ld of %ld\n",
         c1_ok, c1_total, 100.0 * c1_ok / c1_total, cf_ok, 100.0 * cf_ok / c1_total, cf_ok_given_c1, c1_ok);
  printf("C2 holds in %ld of %ld (%.1f%%); closed form correct given C2 %ld of %ld\n", c2_ok, c1_total, 100.0 * c2_ok / c1_total, cf_ok_given_c2, c2_ok);
  for (auto &kv : badhist) printf("  tail winners tied with another winner = %d : %ld selects\n", kv.first, kv.second);
  return 0;
}
