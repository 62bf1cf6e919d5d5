/* oracle/leiden_ref.c — sequential CPU restatement of the Leiden algorithm.  TEST INFRASTRUCTURE ONLY.
 *
 * What the reference runs (src/scanpy/tools/_leiden.py:167-196 of scverse/scanpy @ fabadb94):
 *     leidenalg.find_partition(g, RBConfigurationVertexPartition, weights=..., n_iterations=...,
 *                              resolution_parameter=..., seed=...)            (flavor='leidenalg')
 *     g.community_leiden(objective_function='modularity', weights=..., resolution=..., n_iterations=...)
 * leidenalg (>= 0.10.1) and igraph (>= 0.10.8) are third-party, NOT vendored under /root/reference and not
 * installed, so this file restates the published algorithm (Traag, Waltman & van Eck 2019, "From Louvain
 * to Leiden") in the shape of leidenalg's Optimiser::optimise_partition:
 *     repeat { move_nodes (fast local moving, queue)  ->  merge_nodes_constrained (refinement)
 *              ->  aggregate on the REFINED partition, initial membership = un-refined communities }
 *     until the aggregate stops shrinking;   n_iterations such passes (-1: until a pass changes nothing).
 * Quality = RB configuration model (== Newman modularity with resolution gamma for a symmetric graph):
 *     Q = 1/(2m) * sum_c [ sum_{i,j in c} A_ij  -  gamma * K_c^2 / (2m) ]
 * Label level: the reference's tests contain no Leiden label golden; the labels the reference stored in its
 * in-tree fixture (obs/louvain of pbmc68k_reduced, sc.tl.louvain defaults) are reproduced at ARI 0.94-0.98
 * (tests/test_oracle_leiden_guarantees.py).
 *
 * C ABI (ctypes, see oracle/leiden.py):
 *   int leiden_ref(n, indptr[int64 n+1], indices[int32], weights[double], gamma, n_iterations, seed,
 *                  membership_out[int32 n], modularity_out[double], n_comms_out[int32], passes_out[int32])
 *   int leiden_ref2(..., seed, beta, ...)   beta <= 0: greedy refinement (leidenalg); beta > 0: randomised refinement
 *                  restricted to well-connected vertices / communities (the paper; igraph uses beta = 0.01)
 *   double modularity_ref(n, indptr, indices, weights, gamma, membership)
 * The input must be a symmetric adjacency in CSR (both (i,j) and (j,i) stored); a diagonal entry A_ii
 * counts as sum_{i,j in c} contribution A_ii (this is how aggregated self-loops are stored).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t n;
  int64_t *indptr;
  int32_t *indices;
  double *w;
  double *k;      /* strength incl. self loop */
  double total;   /* 2m = sum k */
} graph_t;

static uint64_t rng_state;
static inline uint64_t rng_next(void) { /* splitmix64 */
  uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static void shuffle(int32_t *a, int32_t n) {
  for (int32_t i = n - 1; i > 0; --i) {
    int32_t j = (int32_t)(rng_next() % (uint64_t)(i + 1));
    int32_t t = a[i]; a[i] = a[j]; a[j] = t;
  }
}

static void graph_strengths(graph_t *g) {
  g->k = (double *)calloc((size_t)g->n, sizeof(double));
  g->total = 0.0;
  for (int32_t i = 0; i < g->n; ++i) {
    double s = 0.0;
    for (int64_t e = g->indptr[i]; e < g->indptr[i + 1]; ++e) s += g->w[e];
    g->k[i] = s;
    g->total += s;
  }
}

/* gain of putting v into community c, up to terms that do not depend on c:
 *   w(v,c) - gamma * k_v * K_c(without v) / (2m)                                   */

/* Fast local moving.  comm[] in/out; K[] community strengths in/out; returns #moves. */
static int64_t move_nodes(const graph_t *g, int32_t *comm, double *K, int32_t *csize, double gamma) {
  const int32_t n = g->n;
  int32_t *queue = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  uint8_t *inq = (uint8_t *)malloc((size_t)n);
  double *cw = (double *)calloc((size_t)n, sizeof(double));
  int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  int32_t *empties = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  int32_t n_empty = 0;
  for (int32_t c = 0; c < n; ++c) if (csize[c] == 0) empties[n_empty++] = c;
  for (int32_t i = 0; i < n; ++i) { queue[i] = i; inq[i] = 1; }
  shuffle(queue, n);
  int64_t head = 0, count = n, moves = 0; /* circular buffer of capacity n */
  const double inv2m = 1.0 / g->total;
  while (count > 0) {
    int32_t v = queue[head]; head = (head + 1) % n; --count; inq[v] = 0;
    const int32_t a = comm[v];
    const double kv = g->k[v];
    int32_t nt = 0;
    for (int64_t e = g->indptr[v]; e < g->indptr[v + 1]; ++e) {
      int32_t u = g->indices[e];
      if (u == v) continue;
      int32_t c = comm[u];
      if (cw[c] == 0.0) touched[nt++] = c;
      cw[c] += g->w[e];
    }
    double best_gain = cw[a] - gamma * kv * (K[a] - kv) * inv2m;
    int32_t best = a;
    for (int32_t t = 0; t < nt; ++t) {
      int32_t c = touched[t];
      if (c == a) continue;
      double gain = cw[c] - gamma * kv * K[c] * inv2m;
      if (gain > best_gain + 1e-15 || (fabs(gain - best_gain) <= 1e-15 && best != a && c < best)) {
        best_gain = gain; best = c;
      }
    }
    /* consider an empty community (only meaningful if v is not alone already) */
    if (csize[a] > 1 && n_empty > 0 && 0.0 > best_gain + 1e-15) { best = empties[n_empty - 1]; best_gain = 0.0; }
    for (int32_t t = 0; t < nt; ++t) cw[touched[t]] = 0.0;
    if (best != a) {
      if (csize[best] == 0) --n_empty;
      K[a] -= kv; K[best] += kv; csize[a]--; csize[best]++;
      if (csize[a] == 0) empties[n_empty++] = a;
      comm[v] = best; ++moves;
      for (int64_t e = g->indptr[v]; e < g->indptr[v + 1]; ++e) {
        int32_t u = g->indices[e];
        if (u != v && comm[u] != best && !inq[u]) { queue[(head + count) % n] = u; ++count; inq[u] = 1; }
      }
    }
  }
  free(queue); free(inq); free(cw); free(touched); free(empties);
  return moves;
}

static double rng_uniform(void) { return (double)(rng_next() >> 11) * (1.0 / 9007199254740992.0); }

/* Refinement: singletons merge into refined communities inside their parent community.
 *   beta <= 0  leidenalg's Optimiser::merge_nodes_constrained as scanpy's flavor='leidenalg' runs it (default
 *              refine_consider_comms = ALL_NEIGH_COMMS): the best non-negative gain wins, greedy, no connectivity test.
 *   beta  > 0  the published algorithm (Traag, Waltman & van Eck 2019, section "Refinement") as igraph's
 *              community_leiden runs it for scanpy's flavor='igraph' (beta = 0.01): only a singleton v that is well
 *              connected to its parent community S,  E(v, S-v) >= gamma k_v (K_S - k_v) / 2m,  moves; only into a refined
 *              community C that is itself well connected,  E(C, S-C) >= gamma K_C (K_S - K_C) / 2m,  with a non-negative
 *              gain; the target is drawn with probability ~ exp(gain / beta). */
static void merge_nodes_constrained(const graph_t *g, const int32_t *parent, int32_t *ref, double gamma, double beta) {
  const int32_t n = g->n;
  double *K = (double *)malloc(sizeof(double) * (size_t)n);
  int32_t *csize = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  double *cw = (double *)calloc((size_t)n, sizeof(double));
  int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  double *Kpar = NULL, *ext = NULL, *prob = NULL;
  if (beta > 0.0) {
    Kpar = (double *)calloc((size_t)n, sizeof(double));   /* strength of each parent community */
    ext = (double *)calloc((size_t)n, sizeof(double));    /* E(C, S-C) per refined community */
    prob = (double *)malloc(sizeof(double) * (size_t)n);
    for (int32_t i = 0; i < n; ++i) {
      Kpar[parent[i]] += g->k[i];
      for (int64_t e = g->indptr[i]; e < g->indptr[i + 1]; ++e) {
        int32_t u = g->indices[e];
        if (u != i && parent[u] == parent[i]) ext[i] += g->w[e];
      }
    }
  }
  for (int32_t i = 0; i < n; ++i) { ref[i] = i; K[i] = g->k[i]; csize[i] = 1; order[i] = i; }
  shuffle(order, n);
  const double inv2m = 1.0 / g->total;
  for (int32_t ii = 0; ii < n; ++ii) {
    int32_t v = order[ii];
    if (csize[ref[v]] != 1) continue; /* only singletons move */
    const double kv = g->k[v];
    if (beta > 0.0 && ext[ref[v]] < gamma * kv * (Kpar[parent[v]] - kv) * inv2m - 1e-15) continue; /* v not well connected */
    int32_t nt = 0;
    for (int64_t e = g->indptr[v]; e < g->indptr[v + 1]; ++e) {
      int32_t u = g->indices[e];
      if (u == v || parent[u] != parent[v]) continue;
      int32_t c = ref[u];
      if (cw[c] == 0.0) touched[nt++] = c;
      cw[c] += g->w[e];
    }
    int32_t best = ref[v];
    if (beta <= 0.0) {
      double best_gain = 0.0; /* staying alone: w=0, K(without v)=0 */
      for (int32_t t = 0; t < nt; ++t) {
        int32_t c = touched[t];
        double gain = cw[c] - gamma * kv * K[c] * inv2m;
        if (gain > best_gain + 1e-15 || (best != ref[v] && fabs(gain - best_gain) <= 1e-15 && c < best)) {
          best_gain = gain; best = c;
        }
      }
    } else {
      /* candidates: staying alone (gain 0) and every well-connected neighbouring refined community with gain >= 0 */
      double max_gain = 0.0, total_p = 0.0;
      int32_t ncand = 0;
      for (int32_t t = 0; t < nt; ++t) {
        int32_t c = touched[t];
        if (c == ref[v]) continue;
        if (ext[c] < gamma * K[c] * (Kpar[parent[v]] - K[c]) * inv2m - 1e-15) continue;
        double gain = cw[c] - gamma * kv * K[c] * inv2m;
        if (gain < 0.0) continue;
        touched[ncand] = c; prob[ncand] = gain; ++ncand;   /* compacts in place: ncand <= t */
        if (gain > max_gain) max_gain = gain;
      }
      /* cw of dropped candidates must still be cleared below: remember them through a second pass over the arcs */
      for (int32_t t = 0; t < ncand; ++t) { prob[t] = exp((prob[t] - max_gain) / beta); total_p += prob[t]; }
      const double p_stay = exp((0.0 - max_gain) / beta);
      total_p += p_stay;
      double r = rng_uniform() * total_p;
      best = ref[v];
      for (int32_t t = 0; t < ncand; ++t) { if (r < prob[t]) { best = touched[t]; break; } r -= prob[t]; }
    }
    const double w_best = (best != ref[v]) ? cw[best] : 0.0;
    for (int64_t e = g->indptr[v]; e < g->indptr[v + 1]; ++e) cw[ref[g->indices[e]]] = 0.0;
    if (best != ref[v]) {
      if (beta > 0.0) { ext[best] += ext[ref[v]] - 2.0 * w_best; ext[ref[v]] = 0.0; }
      K[ref[v]] -= kv; csize[ref[v]]--; K[best] += kv; csize[best]++; ref[v] = best;
    }
  }
  free(K); free(csize); free(order); free(cw); free(touched);
  if (Kpar) { free(Kpar); free(ext); free(prob); }
}

/* relabel labels[] to 0..nc-1 (first-occurrence order); returns nc */
static int32_t compact_labels(int32_t *labels, int32_t n) {
  int32_t *map = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  for (int32_t i = 0; i < n; ++i) map[i] = -1;
  int32_t nc = 0;
  for (int32_t i = 0; i < n; ++i) { if (map[labels[i]] < 0) map[labels[i]] = nc++; labels[i] = map[labels[i]]; }
  free(map);
  return nc;
}

/* aggregate g by labels ref[] (already compact 0..nc-1) */
static graph_t aggregate(const graph_t *g, const int32_t *ref, int32_t nc) {
  graph_t a; a.n = nc;
  /* bucket nodes by community */
  int64_t *start = (int64_t *)calloc((size_t)nc + 1, sizeof(int64_t));
  for (int32_t i = 0; i < g->n; ++i) start[ref[i] + 1]++;
  for (int32_t c = 0; c < nc; ++c) start[c + 1] += start[c];
  int32_t *members = (int32_t *)malloc(sizeof(int32_t) * (size_t)g->n);
  int64_t *pos = (int64_t *)malloc(sizeof(int64_t) * (size_t)nc);
  memcpy(pos, start, sizeof(int64_t) * (size_t)nc);
  for (int32_t i = 0; i < g->n; ++i) members[pos[ref[i]]++] = i;
  double *cw = (double *)calloc((size_t)nc, sizeof(double));
  uint8_t *seen = (uint8_t *)calloc((size_t)nc, 1);
  int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)nc);
  int64_t cap = g->indptr[g->n] + nc + 16, nnz = 0;
  a.indptr = (int64_t *)malloc(sizeof(int64_t) * ((size_t)nc + 1));
  a.indices = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
  a.w = (double *)malloc(sizeof(double) * (size_t)cap);
  a.indptr[0] = 0;
  for (int32_t c = 0; c < nc; ++c) {
    int32_t nt = 0;
    for (int64_t p = start[c]; p < start[c + 1]; ++p) {
      int32_t v = members[p];
      for (int64_t e = g->indptr[v]; e < g->indptr[v + 1]; ++e) {
        int32_t d = ref[g->indices[e]];
        if (!seen[d]) { seen[d] = 1; touched[nt++] = d; }
        cw[d] += g->w[e];
      }
    }
    for (int32_t t = 0; t < nt; ++t) {
      int32_t d = touched[t];
      a.indices[nnz] = d; a.w[nnz] = cw[d]; ++nnz;
      cw[d] = 0.0; seen[d] = 0;
    }
    a.indptr[c + 1] = nnz;
  }
  free(start); free(members); free(pos); free(cw); free(seen); free(touched);
  graph_strengths(&a);
  return a;
}

static void graph_free(graph_t *g) { free(g->indptr); free(g->indices); free(g->w); free(g->k); }

static double quality(const graph_t *g, const int32_t *comm, double gamma) {
  const int32_t n = g->n;
  double *K = (double *)calloc((size_t)n, sizeof(double));
  double in = 0.0;
  for (int32_t i = 0; i < n; ++i) {
    K[comm[i]] += g->k[i];
    for (int64_t e = g->indptr[i]; e < g->indptr[i + 1]; ++e)
      if (comm[g->indices[e]] == comm[i]) in += g->w[e];
  }
  double pen = 0.0;
  for (int32_t c = 0; c < n; ++c) pen += K[c] * K[c];
  free(K);
  return (in - gamma * pen / g->total) / g->total;
}

/* One leidenalg-style optimise_partition pass on the ORIGINAL graph g0, starting from member0[] (values < n).
 * Returns 1 if any node moved. */
static int optimise_pass(const graph_t *g0, int32_t *member0, double gamma, double beta) {
  const int32_t n0 = g0->n;
  graph_t g = *g0; /* current level graph (level 0 borrows g0's arrays) */
  int level0 = 1, improved = 0;
  /* node_of[i] = current-level node that original node i belongs to */
  int32_t *node_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)n0);
  for (int32_t i = 0; i < n0; ++i) node_of[i] = i;
  int32_t *comm = (int32_t *)malloc(sizeof(int32_t) * (size_t)n0);
  memcpy(comm, member0, sizeof(int32_t) * (size_t)n0);
  for (;;) {
    const int32_t n = g.n;
    double *K = (double *)calloc((size_t)n, sizeof(double));
    int32_t *csize = (int32_t *)calloc((size_t)n, sizeof(int32_t));
    for (int32_t i = 0; i < n; ++i) { K[comm[i]] += g.k[i]; csize[comm[i]]++; }
    if (move_nodes(&g, comm, K, csize, gamma) > 0) improved = 1;
    free(K); free(csize);
    /* push the level's communities down to the original nodes */
    for (int32_t i = 0; i < n0; ++i) member0[i] = comm[node_of[i]];
    int32_t *ref = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    merge_nodes_constrained(&g, comm, ref, gamma, beta);
    int32_t nc = compact_labels(ref, n);
    /* number of (non-empty) communities at this level */
    int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    memcpy(tmp, comm, sizeof(int32_t) * (size_t)n);
    int32_t ncomm = compact_labels(tmp, n);
    free(tmp);
    (void)ncomm;
    if (nc >= n) { free(ref); break; } /* refinement merged nothing: the aggregate would not shrink */
    graph_t a = aggregate(&g, ref, nc);
    int32_t *acomm = (int32_t *)malloc(sizeof(int32_t) * (size_t)n0);
    for (int32_t i = 0; i < n; ++i) acomm[ref[i]] = comm[i];
    /* aggregate's community ids must be < a.n: compact them */
    int32_t *amap = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int32_t i = 0; i < n; ++i) amap[i] = -1;
    int32_t nn = 0;
    for (int32_t c = 0; c < nc; ++c) { if (amap[acomm[c]] < 0) amap[acomm[c]] = nn++; acomm[c] = amap[acomm[c]]; }
    for (int32_t i = 0; i < n0; ++i) node_of[i] = ref[node_of[i]];
    free(amap); free(ref);
    if (!level0) graph_free(&g);
    g = a; level0 = 0;
    memcpy(comm, acomm, sizeof(int32_t) * (size_t)nc);
    free(acomm);
  }
  for (int32_t i = 0; i < n0; ++i) member0[i] = comm[node_of[i]];
  if (!level0) graph_free(&g);
  free(node_of); free(comm);
  return improved;
}

/* renumber by decreasing size, ties -> smaller first member */
static int32_t renumber_by_size(int32_t *member, int32_t n) {
  int32_t nc = compact_labels(member, n); /* first-occurrence order == smallest member index order */
  int64_t *size = (int64_t *)calloc((size_t)nc, sizeof(int64_t));
  for (int32_t i = 0; i < n; ++i) size[member[i]]++;
  int32_t *ord = (int32_t *)malloc(sizeof(int32_t) * (size_t)nc);
  for (int32_t c = 0; c < nc; ++c) ord[c] = c;
  /* stable insertion-free sort: simple merge via qsort on (size desc, id asc) */
  for (int32_t i = 1; i < nc; ++i) { /* nc is small (<= few 1e4): binary insertion is fine */
    int32_t x = ord[i]; int32_t lo = 0, hi = i;
    while (lo < hi) { int32_t mid = (lo + hi) / 2; int32_t y = ord[mid];
      if (size[y] > size[x] || (size[y] == size[x] && y < x)) lo = mid + 1; else hi = mid; }
    memmove(ord + lo + 1, ord + lo, sizeof(int32_t) * (size_t)(i - lo)); ord[lo] = x;
  }
  int32_t *newid = (int32_t *)malloc(sizeof(int32_t) * (size_t)nc);
  for (int32_t r = 0; r < nc; ++r) newid[ord[r]] = r;
  for (int32_t i = 0; i < n; ++i) member[i] = newid[member[i]];
  free(size); free(ord); free(newid);
  return nc;
}

int leiden_ref2(int32_t n, const int64_t *indptr, const int32_t *indices, const double *weights, double gamma,
                int32_t n_iterations, uint64_t seed, double beta, int32_t *membership, double *modularity, int32_t *n_comms,
                int32_t *passes) {
  graph_t g; g.n = n; g.indptr = (int64_t *)indptr; g.indices = (int32_t *)indices; g.w = (double *)weights;
  graph_strengths(&g);
  rng_state = seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
  for (int32_t i = 0; i < n; ++i) membership[i] = i;
  int32_t it = 0;
  if (g.total > 0.0) {
    for (;;) {
      int improved = optimise_pass(&g, membership, gamma, beta);
      ++it;
      compact_labels(membership, n);
      if (n_iterations >= 0 ? it >= n_iterations : !improved) break;
    }
  }
  *n_comms = renumber_by_size(membership, n);
  *modularity = g.total > 0.0 ? quality(&g, membership, gamma) : 0.0;
  *passes = it;
  free(g.k);
  return 0;
}

int leiden_ref(int32_t n, const int64_t *indptr, const int32_t *indices, const double *weights, double gamma,
               int32_t n_iterations, uint64_t seed, int32_t *membership, double *modularity, int32_t *n_comms,
               int32_t *passes) {
  return leiden_ref2(n, indptr, indices, weights, gamma, n_iterations, seed, 0.0, membership, modularity, n_comms, passes);
}

double modularity_ref(int32_t n, const int64_t *indptr, const int32_t *indices, const double *weights,
                      double gamma, const int32_t *membership) {
  graph_t g; g.n = n; g.indptr = (int64_t *)indptr; g.indices = (int32_t *)indices; g.w = (double *)weights;
  graph_strengths(&g);
  int32_t *m = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  memcpy(m, membership, sizeof(int32_t) * (size_t)n);
  compact_labels(m, n);
  double q = g.total > 0.0 ? quality(&g, m, gamma) : 0.0;
  free(m); free(g.k);
  return q;
}
