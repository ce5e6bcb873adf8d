// Topology-aware reduction trees for single-process multi-GPU kvstore (MXNET_KVSTORE_USETREE=1).
//
// Reference behaviour: src/kvstore/gpu_topology.h builds, for every root GPU, a binary reduction tree from the
// P2P link matrix (cudaDeviceGetP2PAttribute PerformanceRank) by recursive Kernighan-Lin bipartition, penalising
// links that are already used by a factor 0.7 (gpu_topology.h:269-427, 490-575, 789-960), and CommDeviceTree
// (comm_tree.h) reduces / broadcasts along those trees.
//
// On an NVSwitch box every pair has the same weight and any balanced tree is equally good (the fused P2P kernel
// is the fast path there); this component keeps the reference's behaviour for asymmetric topologies (PCIe
// switches, partial NVLink meshes) and for API parity.  The algorithm below is an independent implementation:
//   tree(S, root): if |S| == 1 stop.  Pick a sub-root among the root's strongest neighbours (bounded backtracking
//   over the top 3), grow two halves A (root) and B (sub-root) by strongest attachment, refine with Kernighan-Lin
//   pair swaps (gain = D_a + D_b - 2 w_ab), recurse into both halves; keep the candidate whose tree has the fewest
//   link-less edges, then the largest total weight.  Edges used by a root's tree get their effective weight
//   multiplied by alpha so that other roots prefer other links.
// Output per root: n-1 edges (level, parent, child); reducing executes levels deepest-first (children add into
// parents), broadcasting runs them in the opposite order.
#include <algorithm>
#include <cmath>
#include <vector>

#include "../c_api.h"

namespace mxb {
namespace topo {

struct Result {
  int misses = 0;        // edges without a direct link (they would be staged through the host)
  float weight = 0.f;    // sum of the (penalised) weights of the edges used
  std::vector<int> edges;
};

struct Builder {
  int n;
  std::vector<float> w;      // effective (penalised) weights, row-major n x n
  static constexpr int kCandidates = 3;
  static constexpr float kRootAffinity = 2.f;

  float W(int a, int b) const { return w[a * n + b]; }

  // Split S into A (contains ra, ceil(|S|/2) nodes) and B (contains rb): grow both sides from their roots, always
  // adding the (node, side) pair with the strongest attachment, then Kernighan-Lin pair swaps that keep ra / rb.
  void split(const std::vector<int>& S, int ra, int rb, bool refine, std::vector<int>* A, std::vector<int>* B) const {
    const int capA = ((int)S.size() + 1) / 2, capB = (int)S.size() - capA;
    std::vector<char> side(n, 0);   // 1 = A, 2 = B
    side[ra] = 1;
    side[rb] = 2;
    int nA = 1, nB = 1;
    while (nA + nB < (int)S.size()) {
      int best = -1, bside = 0;
      float bw = -1.f;
      for (int v : S) {
        if (side[v]) continue;
        // attachment to a side; a direct link to that side's root counts extra: the root needs ~log2(|side|)
        // neighbours inside its own half to reach every level below without leaving the links
        float sa = kRootAffinity * W(v, ra), sb = kRootAffinity * W(v, rb);
        for (int u : S) {
          if (side[u] == 1) sa += W(v, u);
          else if (side[u] == 2) sb += W(v, u);
        }
        if (nA < capA && sa > bw) { bw = sa; best = v; bside = 1; }
        if (nB < capB && sb > bw) { bw = sb; best = v; bside = 2; }
      }
      side[best] = (char)bside;
      (bside == 1 ? nA : nB)++;
    }
    for (int pass = 0; refine && pass < 2 * (int)S.size(); ++pass) {
      float best_gain = 1e-6f;
      int ba = -1, bb = -1;
      for (int a : S) {
        if (side[a] != 1 || a == ra) continue;
        float Da = 0.f;
        for (int u : S)
          if (u != a) Da += side[u] == 1 ? -W(a, u) : W(a, u);
        for (int b : S) {
          if (side[b] != 2 || b == rb) continue;
          float Db = 0.f;
          for (int u : S)
            if (u != b) Db += side[u] == 1 ? W(b, u) : -W(b, u);
          const float gain = Da + Db - 2.f * W(a, b);
          if (gain > best_gain) { best_gain = gain; ba = a; bb = b; }
        }
      }
      if (ba < 0) break;
      side[ba] = 2;
      side[bb] = 1;
    }
    for (int v : S) (side[v] == 1 ? A : B)->push_back(v);
  }

  Result solve(const std::vector<int>& S, int root, int level) const {
    Result best;
    if (S.size() <= 1) return best;
    // candidate sub-roots: the nodes with the strongest link to the root (bounded backtracking over the top few)
    std::vector<int> cand;
    for (int v : S)
      if (v != root) cand.push_back(v);
    std::stable_sort(cand.begin(), cand.end(), [&](int x, int y) { return W(root, x) > W(root, y); });
    const int max_cand = n <= 16 ? n : kCandidates;   // small boxes: full backtracking over the sub-root choice
    if ((int)cand.size() > max_cand) cand.resize(max_cand);
    bool have = false;
    for (int sub : cand) {
      for (int variant = 0; variant < 2; ++variant) {   // partition with and without the Kernighan-Lin refinement
        std::vector<int> A, B;
        split(S, root, sub, variant == 0, &A, &B);
        Result ra = solve(A, root, level + 1), rb = solve(B, sub, level + 1);
        Result r;
        r.misses = ra.misses + rb.misses + (W(root, sub) > 0.f ? 0 : 1);
        r.weight = ra.weight + rb.weight + W(root, sub);
        if (!have || r.misses < best.misses || (r.misses == best.misses && r.weight > best.weight)) {
          r.edges.push_back(level);
          r.edges.push_back(root);
          r.edges.push_back(sub);
          r.edges.insert(r.edges.end(), ra.edges.begin(), ra.edges.end());
          r.edges.insert(r.edges.end(), rb.edges.begin(), rb.edges.end());
          best = std::move(r);
          have = true;
        }
        if (best.misses == 0) break;
      }
      if (best.misses == 0) break;   // an all-linked tree: stop backtracking
    }
    return best;
  }
};

}  // namespace topo
}  // namespace mxb

// W: n x n link weights (0 = no direct P2P link, larger = faster; diagonal ignored).
// out_edges: n roots x (n-1) edges x 3 ints (level, parent, child).
MXB_API int MXBTopoComputeTrees(int n, const float* W, float alpha, int* out_edges) {
  MXB_TRY
  MXB_CHECK(n >= 1 && n <= 64, "topology: 1..64 devices");
  std::vector<float> base(W, W + (size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {   // symmetrise
      const float m = std::max(base[i * n + j], base[j * n + i]);
      base[i * n + j] = base[j * n + i] = m;
    }
  std::vector<float> eff = base;   // penalties persist across roots so different roots prefer different links
  for (int r = 0; r < n; ++r) {
    mxb::topo::Builder b{n, eff};
    std::vector<int> S(n);
    for (int i = 0; i < n; ++i) S[i] = i;
    mxb::topo::Result res = b.solve(S, r, 0);
    MXB_CHECK((int)res.edges.size() == 3 * (n - 1), "topology: internal error (edge count)");
    for (size_t e = 0; e < res.edges.size(); e += 3) {   // a used link is less attractive for the other roots
      const int p = res.edges[e + 1], c = res.edges[e + 2];
      eff[p * n + c] *= alpha;
      eff[c * n + p] *= alpha;
    }
    std::copy(res.edges.begin(), res.edges.end(), out_edges + (size_t)r * (n - 1) * 3);
  }
  MXB_CATCH
}
