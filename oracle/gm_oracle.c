/*
 * gm_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's (chenxuhao/GraphMiner) CPU algorithm for the
 * sorted adjacency-list intersection / difference path and the OpenMP loop nests that
 * drive it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library; the product (libgminer_b200.so) never
 * links, loads or calls it.
 *
 * Parity pin: every solver below is checked (tests/test_oracle.py) against the
 * known-answer tables of the reference READMEs on the bundled citeseer and mico graphs
 * (src/triangle/README.md:51-64, src/clique/README.md:52-64, src/sgl/README.md:51-63,
 * src/motif/README.md:50-60) and against the reference's own unmodified OpenMP binaries
 * (oracle/_ref, built by oracle/build_ref.sh) on generated R-MAT graphs
 * (tests/golden/rmat_counts.json, produced by tools/make_golden.py).
 *
 * Each function cites the reference file:line whose behaviour it restates.  Paths are
 * relative to /root/reference.  vidType=int32, eidType=int64 (include/common.h:35-36).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int32_t vid_t;
typedef int64_t eid_t;

/* A VertexSet view: include/VertexSet.h:21-31 {ptr,set_size,vid}.  Pooled sets have vid=-1
 * (VertexSet.h:32).  */
typedef struct { const vid_t *p; vid_t n; vid_t vid; } vset;

/* ------------------------------------------------------------------------------------ */
/* set operators                                                                        */
/* ------------------------------------------------------------------------------------ */

/* VertexSet::get_intersect_num, VertexSet.h:65-76 (intersection_num :297-299). */
int64_t gmo_intersection_num(const vid_t *a, vid_t na, const vid_t *b, vid_t nb) {
  int64_t num = 0; vid_t i = 0, j = 0;
  while (i < na && j < nb) {
    vid_t l = a[i], r = b[j];
    if (l <= r) i++;
    if (r <= l) j++;
    if (l == r) num++;
  }
  return num;
}

/* VertexSet::operator&, VertexSet.h:53-64 (intersection_set :289-291). */
vid_t gmo_intersection_set(const vid_t *a, vid_t na, const vid_t *b, vid_t nb, vid_t *out) {
  vid_t n = 0, i = 0, j = 0;
  while (i < na && j < nb) {
    vid_t l = a[i], r = b[j];
    if (l <= r) i++;
    if (r <= l) j++;
    if (l == r) out[n++] = l;
  }
  return n;
}

/* VertexSet::intersect_ns(other, upper), VertexSet.h:110-122: stop as soon as EITHER head >= upper. */
int64_t gmo_intersection_num_bound(const vid_t *a, vid_t na, const vid_t *b, vid_t nb, vid_t upper) {
  int64_t num = 0; vid_t i = 0, j = 0;
  while (i < na && j < nb) {
    vid_t l = a[i], r = b[j];
    if (l >= upper) break;
    if (r >= upper) break;
    if (l <= r) i++;
    if (r <= l) j++;
    if (l == r) num++;
  }
  return num;
}

/* VertexSet::intersect(other, upper), VertexSet.h:95-108. */
vid_t gmo_intersection_set_bound(const vid_t *a, vid_t na, const vid_t *b, vid_t nb, vid_t upper, vid_t *out) {
  vid_t n = 0, i = 0, j = 0;
  while (i < na && j < nb) {
    vid_t l = a[i], r = b[j];
    if (l >= upper) break;
    if (r >= upper) break;
    if (l <= r) i++;
    if (r <= l) j++;
    if (l == r) out[n++] = l;
  }
  return n;
}

/* intersect_ns_except (1 or 2 ancestors), VertexSet.h:165-189; pass anc_b = -1 for "one". */
int64_t gmo_intersection_num_except(const vid_t *a, vid_t na, const vid_t *b, vid_t nb, vid_t anc_a, vid_t anc_b) {
  int64_t num = 0; vid_t i = 0, j = 0;
  while (i < na && j < nb) {
    vid_t l = a[i], r = b[j];
    if (l <= r) i++;
    if (r <= l) j++;
    if (l == r && l != anc_a && l != anc_b) num++;
  }
  return num;
}

/* intersect_ns_bound_except, VertexSet.h:150-163. */
int64_t gmo_intersection_num_bound_except(const vid_t *a, vid_t na, const vid_t *b, vid_t nb, vid_t upper, vid_t anc) {
  int64_t num = 0; vid_t i = 0, j = 0;
  while (i < na && j < nb) {
    vid_t l = a[i], r = b[j];
    if (l >= upper) break;
    if (r >= upper) break;
    if (l <= r) i++;
    if (r <= l) j++;
    if (l == r && l != anc) num++;
  }
  return num;
}

/* intersect_except (materialising), VertexSet.h:124-135. */
vid_t gmo_intersection_set_except(const vid_t *a, vid_t na, const vid_t *b, vid_t nb, vid_t anc, vid_t *out) {
  vid_t n = 0, i = 0, j = 0;
  while (i < na && j < nb) {
    vid_t l = a[i], r = b[j];
    if (l <= r) i++;
    if (r <= l) j++;
    if (l == r && l != anc) out[n++] = l;
  }
  return n;
}

/* VertexSet::difference_buf(outBuf, other), src/common/VertexSet.cc:21-43:
 * a \ b, and additionally drops the element equal to other.vid (b_vid). */
vid_t gmo_difference_set(const vid_t *a, vid_t na, const vid_t *b, vid_t nb, vid_t b_vid, vid_t *out) {
  vid_t n = 0, i = 0, j = 0;
  while (i < na && j < nb) {
    vid_t l = a[i], r = b[j];
    if (l <= r) i++;
    if (r <= l) j++;
    if (l < r && l != b_vid) out[n++] = l;
  }
  while (i < na) {
    vid_t l = a[i++];
    if (l != b_vid) out[n++] = l;
  }
  return n;
}

/* VertexSet::difference_buf(outBuf, other, upper), VertexSet.cc:46-67. */
vid_t gmo_difference_set_bound(const vid_t *a, vid_t na, const vid_t *b, vid_t nb, vid_t b_vid, vid_t upper, vid_t *out) {
  vid_t n = 0, i = 0, j = 0;
  while (i < na && j < nb) {
    vid_t l = a[i], r = b[j];
    if (l >= upper) break;
    if (r >= upper) break;
    if (l <= r) i++;
    if (r <= l) j++;
    if (l < r && l != b_vid) out[n++] = l;
  }
  while (i < na) {
    vid_t l = a[i];
    if (l >= upper) break;
    i++;
    if (l != b_vid) out[n++] = l;
  }
  return n;
}

/* difference_num(a,b) = (a-b).size(), VertexSet.h:277-279. */
int64_t gmo_difference_num(const vid_t *a, vid_t na, const vid_t *b, vid_t nb, vid_t b_vid) {
  int64_t n = 0; vid_t i = 0, j = 0;
  while (i < na && j < nb) {
    vid_t l = a[i], r = b[j];
    if (l <= r) i++;
    if (r <= l) j++;
    if (l < r && l != b_vid) n++;
  }
  while (i < na) { if (a[i++] != b_vid) n++; }
  return n;
}

/* VertexSet::difference_ns(other, upper), VertexSet.cc:69-89. */
int64_t gmo_difference_num_bound(const vid_t *a, vid_t na, const vid_t *b, vid_t nb, vid_t b_vid, vid_t upper) {
  int64_t n = 0; vid_t i = 0, j = 0;
  while (i < na && j < nb) {
    vid_t l = a[i], r = b[j];
    if (l >= upper) break;
    if (r >= upper) break;
    if (l <= r) i++;
    if (r <= l) j++;
    if (l < r && l != b_vid) n++;
  }
  while (i < na) {
    vid_t l = a[i];
    if (l >= upper) break;
    i++;
    if (l != b_vid) n++;
  }
  return n;
}

/* VertexSet::bounded(up), VertexSet.h:240-255: length of the prefix with x < up. */
vid_t gmo_bounded(const vid_t *a, vid_t na, vid_t up) {
  if (na > 64) {
    vid_t lo = -1, hi = na;
    while (hi - lo > 1) {
      vid_t mid = (lo + hi) / 2;
      if (a[mid] < up) lo = mid; else hi = mid;
    }
    return lo + 1;
  }
  vid_t i = 0;
  while (i < na && a[i] < up) ++i;
  return i;
}

/* ------------------------------------------------------------------------------------ */
/* graph preparation                                                                    */
/* ------------------------------------------------------------------------------------ */

static inline vset N_(const eid_t *rp, const vid_t *ci, vid_t v) {
  vset s; s.p = ci + rp[v]; s.n = (vid_t)(rp[v + 1] - rp[v]); s.vid = v; return s;
}

/* Graph::orientation, src/common/graph.cc:233-279.  Returns the new edge count; writes the new
 * max degree.  out_colidx must hold at least ne/2 (+slack) entries: caller passes ne. */
eid_t gmo_orient(vid_t nv, const eid_t *rp, const vid_t *ci, eid_t *out_rp, vid_t *out_ci, vid_t *out_maxdeg) {
  vid_t *nd = (vid_t *)calloc((size_t)nv, sizeof(vid_t));
  #pragma omp parallel for schedule(static)
  for (vid_t s = 0; s < nv; s++) {
    vid_t ds = (vid_t)(rp[s + 1] - rp[s]), c = 0;
    for (eid_t e = rp[s]; e < rp[s + 1]; e++) {
      vid_t d = ci[e], dd = (vid_t)(rp[d + 1] - rp[d]);
      if (dd > ds || (dd == ds && d > s)) c++;
    }
    nd[s] = c;
  }
  vid_t md = 0; out_rp[0] = 0;
  for (vid_t s = 0; s < nv; s++) { out_rp[s + 1] = out_rp[s] + nd[s]; if (nd[s] > md) md = nd[s]; }
  #pragma omp parallel for schedule(static)
  for (vid_t s = 0; s < nv; s++) {
    vid_t ds = (vid_t)(rp[s + 1] - rp[s]); eid_t o = out_rp[s];
    for (eid_t e = rp[s]; e < rp[s + 1]; e++) {
      vid_t d = ci[e], dd = (vid_t)(rp[d + 1] - rp[d]);
      if (dd > ds || (dd == ds && d > s)) out_ci[o++] = d;
    }
  }
  free(nd);
  if (out_maxdeg) *out_maxdeg = md;
  return out_rp[nv];
}

/* Graph::init_edgelist(sym_break, ascend=false), graph.cc:297-326.  Returns the number of COO
 * entries written.  Note the reference stops scanning a row at the first u > v ("break"). */
eid_t gmo_edgelist(vid_t nv, const eid_t *rp, const vid_t *ci, int sym_break, vid_t *src, vid_t *dst) {
  eid_t i = 0;
  for (vid_t v = 0; v < nv; v++) {
    for (eid_t e = rp[v]; e < rp[v + 1]; e++) {
      vid_t u = ci[e];
      if (u == v) continue;
      if (sym_break && v < u) break;
      src[i] = v; dst[i] = u; i++;
    }
  }
  return i;
}

/* PartitionedGraph::edgecut_induced_partition1D, src/common/graph_partition.cc:24-132, for ONE part:
 * range [begin,end) plus the 1-hop out-neighbours, order-preserving relabel, vertex-induced CSR.
 * Call once with sub_rp==NULL to size (returns sub_nv, *sub_ne), then again to fill.
 * idx_map[sub_nv] = local->global.  local_begin/local_end as graph_partition.cc:32-33. */
vid_t gmo_partition_part(vid_t nv, const eid_t *rp, const vid_t *ci, vid_t begin, vid_t end,
                         eid_t *sub_rp, vid_t *sub_ci, vid_t *idx_map, eid_t *sub_ne,
                         vid_t *local_begin, vid_t *local_end) {
  int8_t *mask = (int8_t *)calloc((size_t)nv, 1);
  for (vid_t v = begin; v < end; v++) {
    mask[v] = 1;
    for (eid_t e = rp[v]; e < rp[v + 1]; e++) mask[ci[e]] = 1;
  }
  vid_t *newid = (vid_t *)malloc(sizeof(vid_t) * (size_t)nv);
  vid_t m = 0;
  for (vid_t v = 0; v < nv; v++) { newid[v] = m; if (mask[v]) m++; }
  eid_t ne = 0; vid_t k = 0;
  for (vid_t v = 0; v < nv; v++) {
    if (!mask[v]) continue;
    if (idx_map) idx_map[k] = v;
    if (v == begin && local_begin) *local_begin = k;
    if (v == end - 1 && local_end) *local_end = k + 1;
    if (sub_rp) sub_rp[k] = ne;
    for (eid_t e = rp[v]; e < rp[v + 1]; e++) {
      vid_t u = ci[e];
      if (mask[u]) { if (sub_ci) sub_ci[ne] = newid[u]; ne++; }
    }
    k++;
  }
  if (sub_rp) sub_rp[m] = ne;
  if (sub_ne) *sub_ne = ne;
  free(mask); free(newid);
  return m;
}

/* ------------------------------------------------------------------------------------ */
/* solvers (the OpenMP loop nests)                                                      */
/* ------------------------------------------------------------------------------------ */

/* TCSolver, src/triangle/omp_base.cc:15-21, restricted to sources [v_begin, v_end). */
uint64_t gmo_tc_range(vid_t nv, const eid_t *rp, const vid_t *ci, vid_t v_begin, vid_t v_end) {
  uint64_t counter = 0;
  (void)nv;
  #pragma omp parallel for reduction(+ : counter) schedule(dynamic, 1)
  for (vid_t u = v_begin; u < v_end; u++) {
    vset yu = N_(rp, ci, u);
    for (vid_t k = 0; k < yu.n; k++) {
      vset yv = N_(rp, ci, yu.p[k]);
      counter += (uint64_t)gmo_intersection_num(yu.p, yu.n, yv.p, yv.n);
    }
  }
  return counter;
}

uint64_t gmo_tc(vid_t nv, const eid_t *rp, const vid_t *ci) { return gmo_tc_range(nv, rp, ci, 0, nv); }

/* automine_{3,4,5}clique (DAG forms), src/clique/cpu_kernels/automine_omp.h:19-31, 67-83, 138-157;
 * dispatcher :159-183 supports k in {3,4,5} only.  Returns UINT64_MAX for unsupported k.
 * max_deg sizes the per-thread level buffers (VertexSet::MAX_DEGREE). */
uint64_t gmo_kclique_range(vid_t nv, const eid_t *rp, const vid_t *ci, int k, vid_t max_deg,
                           vid_t v_begin, vid_t v_end) {
  uint64_t counter = 0;
  (void)nv;
  if (k < 3 || k > 5) return UINT64_MAX;
  if (max_deg < 1) max_deg = 1;
  #pragma omp parallel reduction(+ : counter)
  {
    vid_t *s1 = (vid_t *)malloc(sizeof(vid_t) * (size_t)max_deg);
    vid_t *s2 = (vid_t *)malloc(sizeof(vid_t) * (size_t)max_deg);
    #pragma omp for schedule(dynamic, 1)
    for (vid_t v0 = v_begin; v0 < v_end; v0++) {
      uint64_t local = 0;
      vset y0 = N_(rp, ci, v0);
      for (vid_t i1 = 0; i1 < y0.n; i1++) {
        vset y1 = N_(rp, ci, y0.p[i1]);
        if (k == 3) { local += (uint64_t)gmo_intersection_num(y0.p, y0.n, y1.p, y1.n); continue; }
        vid_t n1 = gmo_intersection_set(y0.p, y0.n, y1.p, y1.n, s1);
        for (vid_t i2 = 0; i2 < n1; i2++) {
          vset y2 = N_(rp, ci, s1[i2]);
          if (k == 4) { local += (uint64_t)gmo_intersection_num(s1, n1, y2.p, y2.n); continue; }
          vid_t n2 = gmo_intersection_set(s1, n1, y2.p, y2.n, s2);
          for (vid_t i3 = 0; i3 < n2; i3++) {
            vset y3 = N_(rp, ci, s2[i3]);
            local += (uint64_t)gmo_intersection_num(s2, n2, y3.p, y3.n);
          }
        }
      }
      counter += local;
    }
    free(s1); free(s2);
  }
  return counter;
}

uint64_t gmo_kclique(vid_t nv, const eid_t *rp, const vid_t *ci, int k, vid_t max_deg) {
  return gmo_kclique_range(nv, rp, ci, k, max_deg, 0, nv);
}

/* SglSolver patterns, src/sgl/omp_base.cc:5-58 with cpu_kernels/{diamond,rectangle,house,pentagon}.h.
 * pattern: 0 diamond, 1 rectangle, 2 house, 3 pentagon.  Undirected (symmetric) CSR. */
uint64_t gmo_sgl_range(vid_t nv, const eid_t *rp, const vid_t *ci, int pattern, vid_t max_deg,
                       vid_t v_begin, vid_t v_end) {
  uint64_t counter = 0;
  (void)nv;
  if (pattern < 0 || pattern > 3) return UINT64_MAX;
  if (max_deg < 1) max_deg = 1;
  #pragma omp parallel reduction(+ : counter)
  {
    vid_t *s1 = (vid_t *)malloc(sizeof(vid_t) * (size_t)max_deg);
    #pragma omp for schedule(dynamic, 1)
    for (vid_t v0 = v_begin; v0 < v_end; v0++) {
      vset y0 = N_(rp, ci, v0);
      for (vid_t i1 = 0; i1 < y0.n; i1++) {
        vid_t v1 = y0.p[i1];
        if (v1 >= v0) break;
        vset y1 = N_(rp, ci, v1);
        if (pattern == 0) {                       /* diamond.h:1-14 */
          vid_t n = gmo_intersection_set(y0.p, y0.n, y1.p, y1.n, s1);
          for (vid_t a = 0; a < n; a++)
            for (vid_t b = 0; b < n; b++) { if (s1[b] >= s1[a]) break; counter += 1; }
        } else if (pattern == 1) {                /* rectangle.h:1-11 */
          for (vid_t i2 = 0; i2 < y0.n; i2++) {
            vid_t v2 = y0.p[i2];
            if (v2 >= v1) break;
            vset y2 = N_(rp, ci, v2);
            counter += (uint64_t)gmo_intersection_num_bound(y1.p, y1.n, y2.p, y2.n, v0);
          }
        } else if (pattern == 2) {                /* house.h:1-17 */
          vid_t n = gmo_intersection_set(y0.p, y0.n, y1.p, y1.n, s1);
          for (vid_t a = 0; a < n; a++) {
            vid_t v2 = s1[a];
            for (vid_t i3 = 0; i3 < y1.n; i3++) {
              vid_t v3 = y1.p[i3];
              if (v3 == v0 || v3 == v2) continue;
              vset y3 = N_(rp, ci, v3);
              counter += (uint64_t)gmo_intersection_num_except(y0.p, y0.n, y3.p, y3.n, v1, v2);
            }
          }
        } else {                                  /* pentagon.h:1-17 */
          for (vid_t i2 = 0; i2 < y0.n; i2++) {
            vid_t v2 = y0.p[i2];
            if (v2 >= v1) break;
            vset y2 = N_(rp, ci, v2);
            for (vid_t i3 = 0; i3 < y2.n; i3++) {
              vid_t v3 = y2.p[i3];
              if (v3 >= v0) break;
              if (v3 == v1) continue;
              vset y3 = N_(rp, ci, v3);
              counter += (uint64_t)gmo_intersection_num_bound_except(y1.p, y1.n, y3.p, y3.n, v0, v2);
            }
          }
        }
      }
    }
    free(s1);
  }
  return counter;
}

uint64_t gmo_sgl(vid_t nv, const eid_t *rp, const vid_t *ci, int pattern, vid_t max_deg) {
  return gmo_sgl_range(nv, rp, ci, pattern, max_deg, 0, nv);
}

/* automine_3motif, src/motif/cpu_kernels/automine_base.h:2-22.  out[0]=wedges, out[1]=triangles. */
static void motif3_range(const eid_t *rp, const vid_t *ci, vid_t v_begin, vid_t v_end, uint64_t *out) {
  uint64_t c0 = 0, c1 = 0;
  #pragma omp parallel for schedule(dynamic, 1) reduction(+ : c0, c1)
  for (vid_t v0 = v_begin; v0 < v_end; v0++) {
    vset y0 = N_(rp, ci, v0);
    vid_t f0 = gmo_bounded(y0.p, y0.n, v0);
    for (vid_t i1 = 0; i1 < y0.n; i1++) {
      vid_t v1 = y0.p[i1]; vset y1 = N_(rp, ci, v1);
      c0 += (uint64_t)gmo_difference_num_bound(y0.p, y0.n, y1.p, y1.n, v1, v1);
    }
    for (vid_t i1 = 0; i1 < f0; i1++) {
      vid_t v1 = y0.p[i1]; vset y1 = N_(rp, ci, v1);
      c1 += (uint64_t)gmo_intersection_num_bound(y0.p, f0, y1.p, y1.n, v1);
    }
  }
  out[0] += c0; out[1] += c1;
}

/* automine_4motif, automine_base.h:24-75.  Index order 0=3-star 1=4-path 2=tailed-triangle
 * 3=4-cycle 4=diamond 5=4-clique.  vid bookkeeping follows VertexSet: CSR rows carry their vertex id,
 * pooled results carry -1; difference drops `other.vid` (VertexSet.cc:29,37). */
static void motif4_range(const eid_t *rp, const vid_t *ci, vid_t max_deg, vid_t v_begin, vid_t v_end, uint64_t *out) {
  uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
  if (max_deg < 1) max_deg = 1;
  #pragma omp parallel reduction(+ : c0, c1, c2, c3, c4, c5)
  {
    size_t md = (size_t)max_deg;
    vid_t *buf = (vid_t *)malloc(sizeof(vid_t) * md * 8);
    vid_t *y0n1f1 = buf, *y0y1 = buf + md, *y0f0y1f1 = buf + 2 * md, *n0y1 = buf + 3 * md,
          *y0n1 = buf + 4 * md, *y0f0n1f1 = buf + 5 * md, *tmp = buf + 6 * md;
    #pragma omp for schedule(dynamic, 1)
    for (vid_t v0 = v_begin; v0 < v_end; v0++) {
      vset y0 = N_(rp, ci, v0);
      vid_t f0 = gmo_bounded(y0.p, y0.n, v0);
      for (vid_t i1 = 0; i1 < y0.n; i1++) {
        vid_t v1 = y0.p[i1]; vset y1 = N_(rp, ci, v1);
        vid_t n = gmo_difference_set_bound(y0.p, y0.n, y1.p, y1.n, v1, v1, y0n1f1);
        for (vid_t i2 = 0; i2 < n; i2++) {
          vid_t v2 = y0n1f1[i2]; vset y2 = N_(rp, ci, v2);
          c0 += (uint64_t)gmo_difference_num_bound(y0n1f1, n, y2.p, y2.n, v2, v2);
        }
      }
      for (vid_t i1 = 0; i1 < f0; i1++) {
        vid_t v1 = y0.p[i1]; vset y1 = N_(rp, ci, v1);
        vid_t n_y0y1 = gmo_intersection_set(y0.p, y0.n, y1.p, y1.n, y0y1);
        vid_t n_y0f0y1f1 = gmo_intersection_set_bound(y0.p, f0, y1.p, y1.n, v1, y0f0y1f1);
        vid_t n_n0y1 = gmo_difference_set(y1.p, y1.n, y0.p, y0.n, v0, n0y1);   /* also n0f0y1 (automine_base.h:47-48) */
        vid_t n_y0n1 = gmo_difference_set(y0.p, y0.n, y1.p, y1.n, v1, y0n1);
        vid_t n_y0f0n1f1 = gmo_difference_set_bound(y0.p, f0, y1.p, y1.n, v1, v1, y0f0n1f1);
        for (vid_t i2 = 0; i2 < n_y0y1; i2++) {
          vid_t v2 = y0y1[i2]; vset y2 = N_(rp, ci, v2);
          c4 += (uint64_t)gmo_difference_num_bound(y0y1, n_y0y1, y2.p, y2.n, v2, v2);
          vid_t nt = gmo_difference_set(y2.p, y2.n, y0.p, y0.n, v0, tmp);     /* n0y2, pooled => vid -1 */
          c2 += (uint64_t)gmo_difference_num(tmp, nt, y1.p, y1.n, v1);
        }
        for (vid_t i2 = 0; i2 < n_y0f0y1f1; i2++) {
          vid_t v2 = y0f0y1f1[i2]; vset y2 = N_(rp, ci, v2);
          c5 += (uint64_t)gmo_intersection_num_bound(y0f0y1f1, n_y0f0y1f1, y2.p, y2.n, v2);
        }
        for (vid_t i2 = 0; i2 < n_y0n1; i2++) {
          vid_t v2 = y0n1[i2]; vset y2 = N_(rp, ci, v2);
          c1 += (uint64_t)gmo_difference_num(n0y1, n_n0y1, y2.p, y2.n, v2);
        }
        for (vid_t i2 = 0; i2 < n_y0f0n1f1; i2++) {
          vid_t v2 = y0f0n1f1[i2]; vset y2 = N_(rp, ci, v2);
          c3 += (uint64_t)gmo_intersection_num_bound(n0y1, n_n0y1, y2.p, y2.n, v0);
        }
      }
    }
    free(buf);
  }
  out[0] += c0; out[1] += c1; out[2] += c2; out[3] += c3; out[4] += c4; out[5] += c5;
}

/* MotifSolver (omp_base), src/motif/omp_base.cc:8-31.  k in {3,4}; out has 2 or 6 entries (zeroed here). */
int gmo_motif_range(vid_t nv, const eid_t *rp, const vid_t *ci, int k, vid_t max_deg,
                    vid_t v_begin, vid_t v_end, uint64_t *out) {
  (void)nv;
  if (k == 3) { out[0] = out[1] = 0; motif3_range(rp, ci, v_begin, v_end, out); return 0; }
  if (k == 4) { memset(out, 0, 6 * sizeof(uint64_t)); motif4_range(rp, ci, max_deg, v_begin, v_end, out); return 0; }
  return -1;
}

int gmo_motif(vid_t nv, const eid_t *rp, const vid_t *ci, int k, vid_t max_deg, uint64_t *out) {
  return gmo_motif_range(nv, rp, ci, k, max_deg, 0, nv, out);
}

/* MotifSolver (omp_formula), src/motif/omp_formula.cc:8-52 + cpu_kernels/automine_formula.h:2-56:
 * closed forms from the per-edge triangle count; only 4-cycle and 4-clique are enumerated. */
int gmo_motif_formula(vid_t nv, const eid_t *rp, const vid_t *ci, int k, vid_t max_deg, uint64_t *out) {
  if (max_deg < 1) max_deg = 1;
  if (k == 3) {
    uint64_t c0 = 0, c1 = 0;
    #pragma omp parallel for schedule(dynamic, 1) reduction(+ : c0, c1)
    for (vid_t v0 = 0; v0 < nv; v0++) {
      vset y0 = N_(rp, ci, v0);
      uint64_t n = (uint64_t)y0.n;
      c0 += n * (n - 1);
      vid_t f0 = gmo_bounded(y0.p, y0.n, v0);
      for (vid_t i1 = 0; i1 < f0; i1++) {
        vid_t v1 = y0.p[i1]; vset y1 = N_(rp, ci, v1);
        c1 += (uint64_t)gmo_intersection_num_bound(y0.p, y0.n, y1.p, y1.n, v1);
      }
    }
    out[0] = c0 / 2 - 3 * c1; out[1] = c1;
    return 0;
  }
  if (k != 4) return -1;
  uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
  #pragma omp parallel reduction(+ : c0, c1, c2, c3, c4, c5)
  {
    size_t md = (size_t)max_deg;
    vid_t *buf = (vid_t *)malloc(sizeof(vid_t) * md * 3);
    vid_t *y0f0y1f1 = buf, *n0f0y1 = buf + md, *y0f0n1f1 = buf + 2 * md;
    #pragma omp for schedule(dynamic, 1)
    for (vid_t v0 = 0; v0 < nv; v0++) {
      vset y0 = N_(rp, ci, v0);
      vid_t f0 = gmo_bounded(y0.p, y0.n, v0);
      for (vid_t i1 = 0; i1 < f0; i1++) {
        vid_t v1 = y0.p[i1]; vset y1 = N_(rp, ci, v1);
        uint64_t tri = (uint64_t)gmo_intersection_num(y0.p, y0.n, y1.p, y1.n);
        c4 += tri * (tri - 1);
        uint64_t staru = (uint64_t)y0.n - tri - 1, starv = (uint64_t)y1.n - tri - 1;
        c2 += tri * (staru + starv);
        c1 += staru * starv;
        c0 += staru * (staru - 1);
        c0 += starv * (starv - 1);
        vid_t na = gmo_intersection_set_bound(y0.p, f0, y1.p, y1.n, v1, y0f0y1f1);
        vid_t nb = gmo_difference_set(y1.p, y1.n, y0.p, y0.n, v0, n0f0y1);
        vid_t nc = gmo_difference_set_bound(y0.p, f0, y1.p, y1.n, v1, v1, y0f0n1f1);
        for (vid_t i2 = 0; i2 < na; i2++) {
          vset y2 = N_(rp, ci, y0f0y1f1[i2]);
          c5 += (uint64_t)gmo_intersection_num_bound(y0f0y1f1, na, y2.p, y2.n, y0f0y1f1[i2]);
        }
        for (vid_t i2 = 0; i2 < nc; i2++) {
          vset y2 = N_(rp, ci, y0f0n1f1[i2]);
          c3 += (uint64_t)gmo_intersection_num_bound(n0f0y1, nb, y2.p, y2.n, v0);
        }
      }
    }
    free(buf);
  }
  uint64_t t[6] = {c0, c1, c2, c3, c4, c5};
  t[4] = t[4] / 2 - t[5] * 6;          /* omp_formula.cc:42-45 */
  t[2] = t[2] / 2 - t[4] * 2;
  t[1] = t[1] - t[3] * 4;
  t[0] = t[0] / 6 - t[2] / 3;
  memcpy(out, t, sizeof(t));
  return 0;
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline must still use the whole host */
void gmo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int gmo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
