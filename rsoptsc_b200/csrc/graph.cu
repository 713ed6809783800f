// Dense / graph pieces of RepresentationMap (R/RepresentationMap.R:37-39, 57-66, 89), the stage
// immediately downstream of the similarity matrix:
//   dist_flat   stats::dist(flat_embedding) as a full n x n matrix                    (:57)
//   adjacency   adj[S > 0] = 1, diag(adj) = 0                                         (:59-62)
//   components  igraph::components(graph_from_adjacency_matrix(adj, "undirected"))    (:64-66)
//   distances   igraph::distances(graph): all-pairs unweighted shortest paths         (:89)
// Component joining (JoinGraphComponents) stays in R between components() and distances().
#include <algorithm>

#include "internal.h"
#include "sparse.h"

namespace rs {

static constexpr int GT = 256;

// ---- dist_flat: HBM write-bound (8 n^2 B out, 8 n dim B in) -------------------------------------
// R's dist() accumulates dev*dev over the columns left to right and takes sqrt: reproduced without FMA.
__global__ void k_dist_flat(const double* __restrict__ emb, int64_t n, int dim, double* __restrict__ D) {
  const int64_t j = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * GT + threadIdx.x; i < n; i += (int64_t)gridDim.x * GT) {
    double acc = 0;
    for (int c = 0; c < dim; ++c) {
      const double dev = __dsub_rn(emb[i + (int64_t)c * n], emb[j + (int64_t)c * n]);
      acc = __dadd_rn(acc, __dmul_rn(dev, dev));
    }
    D[i + j * n] = sqrt(acc);
  }
}
void dist_flat(const double* d_emb, int64_t n, int dim, double* d_D) {
  Phase ph("dist_flat");
  RS_REQUIRE(n <= 65535, "dist_flat: n too large for this launch shape");
  dim3 grid((unsigned)std::min<int64_t>(cdiv(n, GT), 64), (unsigned)n);
  RS_LAUNCH(k_dist_flat, grid, GT, 0, d_emb, n, dim, d_D);
}

// ---- adjacency -------------------------------------------------------------------------------------
// HBM-bound: 8 n^2 B in, 8 n^2 B out; one column per blockIdx.y (no 64-bit div/mod per element)
__global__ void k_adjacency(const double* __restrict__ S, int64_t n, double* __restrict__ A) {
  const int64_t j = blockIdx.y;
  const double* src = S + j * n;
  double* dst = A + j * n;
  for (int64_t i = (int64_t)blockIdx.x * GT + threadIdx.x; i < n; i += (int64_t)gridDim.x * GT)
    dst[i] = (src[i] > 0.0 && i != j) ? 1.0 : 0.0;
}
void adjacency(const double* d_S, int64_t n, double* d_A) {
  Phase ph("adjacency");
  RS_REQUIRE(n <= 65535, "adjacency: n too large for this launch shape");
  dim3 grid((unsigned)std::min<int64_t>(cdiv(n, GT), 8), (unsigned)n);
  RS_LAUNCH(k_adjacency, grid, GT, 0, d_S, n, d_A);
}

// ---- undirected neighbour lists from a dense adjacency: edge {i,j} if adj[i,j] != 0 or adj[j,i] != 0
struct Graph {
  Compressed byrow, bycol;  // neighbours(v) = byrow line v  U  bycol line v
};
static void build_graph(const double* d_adj, int64_t n, Graph& g) {
  compress(d_adj, n, true, g.byrow);
  compress(d_adj, n, false, g.bycol);
}

// ---- connected components: min-label propagation ------------------------------------------------------
__global__ void k_label_init(int* __restrict__ lab, int64_t n) {
  const int64_t v = (int64_t)blockIdx.x * GT + threadIdx.x;
  if (v < n) lab[v] = (int)v;
}
__global__ void k_label_step(const int* __restrict__ p1, const int* __restrict__ i1, const int* __restrict__ p2,
                             const int* __restrict__ i2, int64_t n, int* __restrict__ lab, int* __restrict__ changed) {
  const int64_t v = (int64_t)blockIdx.x * GT + threadIdx.x;
  if (v >= n) return;
  int m = lab[v];
  for (int e = p1[v]; e < p1[v + 1]; ++e) m = min(m, lab[i1[e]]);
  for (int e = p2[v]; e < p2[v + 1]; ++e) m = min(m, lab[i2[e]]);
  // pointer jumping: the label of my label is at least as small
  m = min(m, lab[m]);
  if (m < lab[v]) { lab[v] = m; *changed = 1; }
}
void graph_components(const double* d_adj, int64_t n, int32_t* h_membership, int32_t* n_components) {
  Phase ph("components");
  cudaStream_t s = ctx().stream;
  Graph g;
  build_graph(d_adj, n, g);
  DevBuf<int> lab(n), changed(1);
  RS_LAUNCH(k_label_init, (unsigned)cdiv(n, GT), GT, 0, lab.p, n);
  for (int64_t iter = 0; iter <= n; ++iter) {
    changed.zero(s);
    RS_LAUNCH(k_label_step, (unsigned)cdiv(n, GT), GT, 0, g.byrow.ptr.p, g.byrow.idx.p, g.bycol.ptr.p, g.bycol.idx.p, n,
              lab.p, changed.p);
    int hc = 0;
    RS_CUDA(cudaMemcpyAsync(&hc, changed.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    RS_CUDA(cudaStreamSynchronize(s));
    if (!hc) break;
  }
  std::vector<int> h(n);
  lab.download(h.data(), n, s);
  RS_CUDA(cudaStreamSynchronize(s));
  // igraph numbers components in order of their lowest vertex (1-based)
  std::vector<int> id(n, 0);
  int next = 0;
  for (int64_t v = 0; v < n; ++v) {
    const int root = h[v];  // converged label == smallest vertex of the component
    if (id[root] == 0) id[root] = ++next;
    if (h_membership) h_membership[v] = id[root];
  }
  if (n_components) *n_components = next;
}

// ---- all-pairs unweighted shortest paths: one CTA per source, level-synchronous BFS in shared memory ----
__global__ void k_bfs_all(const int* __restrict__ p1, const int* __restrict__ i1, const int* __restrict__ p2,
                          const int* __restrict__ i2, int64_t n, int64_t src0, double* __restrict__ D) {
  extern __shared__ unsigned short lvl[];  // n entries, 0xFFFF = unreached
  __shared__ int s_changed;
  const int64_t src = src0 + blockIdx.x;
  if (src >= n) return;
  for (int64_t v = threadIdx.x; v < n; v += GT) lvl[v] = 0xFFFF;
  __syncthreads();
  if (threadIdx.x == 0) lvl[src] = 0;
  __syncthreads();
  for (int level = 0; level < 0xFFFE; ++level) {
    if (threadIdx.x == 0) s_changed = 0;
    __syncthreads();
    for (int64_t v = threadIdx.x; v < n; v += GT) {
      if (lvl[v] != level) continue;
      for (int e = p1[v]; e < p1[v + 1]; ++e) {
        const int u = i1[e];
        if (lvl[u] == 0xFFFF) { lvl[u] = (unsigned short)(level + 1); s_changed = 1; }  // benign race: same value
      }
      for (int e = p2[v]; e < p2[v + 1]; ++e) {
        const int u = i2[e];
        if (lvl[u] == 0xFFFF) { lvl[u] = (unsigned short)(level + 1); s_changed = 1; }
      }
    }
    __syncthreads();
    if (!s_changed) break;
    __syncthreads();
  }
  double* col = D + src * n;  // symmetric: column src == row src
  for (int64_t v = threadIdx.x; v < n; v += GT) col[v] = lvl[v] == 0xFFFF ? INFINITY : (double)lvl[v];
}
void graph_distances(const double* d_adj, int64_t n, double* d_D) {
  Phase ph("bfs");
  RS_REQUIRE(n < 65535, "graph_distances: more than 65534 vertices");
  Graph g;
  build_graph(d_adj, n, g);
  const size_t smem = sizeof(unsigned short) * (size_t)n;
  RS_REQUIRE(smem <= 200 * 1024, "graph_distances: level array does not fit in shared memory");
  if (smem > 48 * 1024) RS_CUDA(cudaFuncSetAttribute(k_bfs_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int64_t s0 = 0; s0 < n; s0 += 65535) {
    const unsigned grid = (unsigned)std::min<int64_t>(65535, n - s0);
    RS_LAUNCH(k_bfs_all, grid, GT, smem, g.byrow.ptr.p, g.byrow.idx.p, g.bycol.ptr.p, g.bycol.idx.p, n, s0, d_D);
  }
  RS_CUDA(cudaStreamSynchronize(ctx().stream));
}

}  // namespace rs
