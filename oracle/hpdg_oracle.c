/*
 * hpdg_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference algorithms on the north-star path of
 * c1887/dune-hpdg.  See hpdg_oracle.h for the rules on who may call this and for
 * the parity pin status.  Every function cites the reference file:line it follows
 * (paths relative to /root/reference/dune/hpdg/).
 *
 * Deliberately NOT sum-factorised: the matrix-free apply and the assemblers are the
 * reference's quadrature-point loops over all n_e shape functions (with the 1-D
 * factor tables cached per degree, as the reference's AssemblyCache does,
 * localfunctions/assemblycache.hh:43-77), so this is an independent formulation
 * from the CUDA product path.
 */
#include "hpdg_oracle.h"
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXP 13 /* functionspacebases/dynamicqknode.hh:84: orders 0..13 */
#define MAXN (MAXP + 1)

/* ------------------------------------------------------------------------ */
/* 1-D building blocks                                                       */
/* ------------------------------------------------------------------------ */

static void legendre(int n, double x, double* P, double* dP) {
  /* P_n(x), P_n'(x) on [-1,1] by the three-term recurrence */
  double p0 = 1.0, p1 = x, d0 = 0.0, d1 = 1.0;
  if (n == 0) { *P = 1.0; *dP = 0.0; return; }
  for (int k = 2; k <= n; k++) {
    double p2 = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
    double d2 = d0 + (2.0 * k - 1.0) * p1;
    p0 = p1; p1 = p2; d0 = d1; d1 = d2;
  }
  *P = p1; *dP = d1;
}

/* Gauss-Lobatto nodes of the (p+1)-point rule mapped to [0,1], ascending.
 * qkgllocalbasis.hh:222-234 takes them from dune-geometry's GaussLobatto rule of order
 * 2p-1 (p+1 points) and sorts them; the table itself is un-vendored, so the nodes are
 * recomputed here as the roots of P_p' plus the end points (Newton, to machine eps). */
void orc_gl_nodes(int p, double* x) {
  if (p == 0) { x[0] = 0.5; return; } /* Q0: single midpoint dof (qkgllocalbasis.hh:224 returns early) */
  x[0] = 0.0; x[p] = 1.0;
  for (int i = 1; i < p; i++) {
    /* Chebyshev-Gauss-Lobatto initial guess, then Newton on q(t) = P_p'(t):
     * q' = P_p'' = (2 t P_p' - p(p+1) P_p)/(1-t^2) */
    double t = -cos(M_PI * i / p);
    for (int it = 0; it < 100; it++) {
      double P, dP; legendre(p, t, &P, &dP);
      double ddP = (2.0 * t * dP - (double)p * (p + 1.0) * P) / (1.0 - t * t);
      double dt = dP / ddP;
      t -= dt;
      if (fabs(dt) < 1e-17) break;
    }
    x[i] = 0.5 * (t + 1.0);
  }
  /* symmetrise (the tabulated rules are symmetric) */
  for (int i = 0; i <= p / 2; i++) {
    double a = 0.5 * (x[i] + (1.0 - x[p - i]));
    x[i] = a; x[p - i] = 1.0 - a;
  }
}

/* m-point Gauss-Legendre rule on [0,1] (exact to degree 2m-1); stands in for
 * dune-geometry's QuadratureRules<double,1>::rule(cube, order) reached through
 * dune-fufem's QuadratureRuleCache (ipdgoperator.hh:133-134,251-253). */
void orc_gauss_legendre(int m, double* x, double* w) {
  for (int i = 0; i < m; i++) {
    double t = -cos(M_PI * (i + 0.75) / (m + 0.5));
    double P, dP;
    for (int it = 0; it < 100; it++) {
      legendre(m, t, &P, &dP);
      double dt = P / dP;
      t -= dt;
      if (fabs(dt) < 1e-17) break;
    }
    legendre(m, t, &P, &dP);
    x[i] = 0.5 * (t + 1.0);
    w[i] = 1.0 / ((1.0 - t * t) * dP * dP); /* = 0.5 * 2/((1-t^2) P'^2) */
  }
}

/* qkgllocalbasis.hh:43-50 */
double orc_lagrange(int p, const double* nodes, int i, double x) {
  double r = 1.0;
  for (int j = 0; j <= p; j++)
    if (j != i) r *= (x - nodes[j]) / (nodes[i] - nodes[j]);
  return r;
}

/* qkgllocalbasis.hh:53-67 */
double orc_lagrange_prime(int p, const double* nodes, int i, double x) {
  double r = 0.0;
  for (int j = 0; j <= p; j++)
    if (j != i) {
      double prod = 1.0 / (nodes[i] - nodes[j]);
      for (int l = 0; l <= p; l++)
        if (l != i && l != j) prod *= (x - nodes[l]) / (nodes[i] - nodes[l]);
      r += prod;
    }
  return r;
}

/* Per-degree caches (the analogue of QkGLVaryingOrderCache, lagrange/qkcache.hh:30-76, and
 * of AssemblyCache).  tab[p][m] : values/derivatives of the p+1 1-D shape functions at the
 * m Gauss points; end[p] : at the two end points 0 and 1. */
typedef struct {
  int ready;
  double nodes[MAXN];
  double endv[2][MAXN], endd[2][MAXN];
} deg_tab;
typedef struct {
  int ready;
  double x[MAXN + 1], w[MAXN + 1];
} gauss_tab;
typedef struct {
  int ready;
  double v[(MAXN + 1) * MAXN], d[(MAXN + 1) * MAXN]; /* [q][i] */
} eval_tab;

static deg_tab g_deg[MAXP + 1];
static gauss_tab g_gauss[MAXN + 2];
static eval_tab* g_eval[MAXP + 1][MAXN + 2];

static const deg_tab* get_deg(int p) {
  deg_tab* t = &g_deg[p];
  if (!t->ready) {
#pragma omp critical(orc_cache)
    if (!t->ready) {
      orc_gl_nodes(p, t->nodes);
      for (int s = 0; s < 2; s++)
        for (int i = 0; i <= p; i++) {
          t->endv[s][i] = orc_lagrange(p, t->nodes, i, (double)s);
          t->endd[s][i] = orc_lagrange_prime(p, t->nodes, i, (double)s);
        }
      t->ready = 1;
    }
  }
  return t;
}
static const gauss_tab* get_gauss(int m) {
  gauss_tab* t = &g_gauss[m];
  if (!t->ready) {
#pragma omp critical(orc_cache)
    if (!t->ready) { orc_gauss_legendre(m, t->x, t->w); t->ready = 1; }
  }
  return t;
}
static const eval_tab* get_eval(int p, int m) {
  if (!g_eval[p][m]) {
    const deg_tab* dt = get_deg(p);
    const gauss_tab* gt = get_gauss(m);
#pragma omp critical(orc_cache)
    if (!g_eval[p][m]) {
      eval_tab* t = (eval_tab*)calloc(1, sizeof(eval_tab));
      for (int q = 0; q < m; q++)
        for (int i = 0; i <= p; i++) {
          t->v[q * MAXN + i] = orc_lagrange(p, dt->nodes, i, gt->x[q]);
          t->d[q * MAXN + i] = orc_lagrange_prime(p, dt->nodes, i, gt->x[q]);
        }
      t->ready = 1;
      g_eval[p][m] = t;
    }
  }
  return g_eval[p][m];
}

/* ------------------------------------------------------------------------ */
/* mesh / basis                                                              */
/* ------------------------------------------------------------------------ */
struct omesh {
  int dim;
  int n[3];
  double L[3], h[3];
  long nelem;
  int* deg;   /* per element degree (dynamicdgqkglbasis.hh:46,133-140) */
  int* pdeg;  /* degree entering the face penalty (== deg on the finest level; on Galerkin
                 coarse levels the fine degrees, ordertransfer.hh:124-144) */
  long* off;  /* block offsets: dynamicbvector.hh:366-373 */
  double sigma;
  int dirichlet;
};

static int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

omesh* orc_mesh_create(int dim, const int* n, const double* L, const int* degree,
                       const int* pen_degree, double sigma, int dirichlet) {
  omesh* m = (omesh*)calloc(1, sizeof(omesh));
  m->dim = dim;
  m->nelem = 1;
  for (int d = 0; d < 3; d++) {
    m->n[d] = d < dim ? n[d] : 1;
    m->L[d] = d < dim ? L[d] : 1.0;
    m->h[d] = m->L[d] / m->n[d];
    m->nelem *= m->n[d];
  }
  m->deg = (int*)malloc(sizeof(int) * m->nelem);
  m->pdeg = (int*)malloc(sizeof(int) * m->nelem);
  m->off = (long*)malloc(sizeof(long) * (m->nelem + 1));
  m->off[0] = 0;
  for (long e = 0; e < m->nelem; e++) {
    m->deg[e] = degree[e];
    m->pdeg[e] = pen_degree ? pen_degree[e] : degree[e];
    m->off[e + 1] = m->off[e] + ipow(degree[e] + 1, dim); /* dynamicdgqkglbasis.hh:104-121 */
  }
  m->sigma = sigma;
  m->dirichlet = dirichlet;
  return m;
}
void orc_mesh_destroy(omesh* m) {
  if (!m) return;
  free(m->deg); free(m->pdeg); free(m->off); free(m);
}
long orc_mesh_dimension(const omesh* m) { return m->off[m->nelem]; }
long orc_mesh_nelem(const omesh* m) { return m->nelem; }
void orc_mesh_offsets(const omesh* m, long* off) { memcpy(off, m->off, sizeof(long) * (m->nelem + 1)); }

static void elem_ijk(const omesh* m, long e, int* ijk) {
  ijk[0] = (int)(e % m->n[0]);
  ijk[1] = (int)((e / m->n[0]) % m->n[1]);
  ijk[2] = (int)(e / ((long)m->n[0] * m->n[1]));
}
/* face f of the reference cube: 0/1 = x-/x+, 2/3 = y-/y+, 4/5 = z-/z+.  Returns neighbour
 * element index or -1 on the domain boundary. */
static long neighbour(const omesh* m, long e, int f) {
  int ijk[3]; elem_ijk(m, e, ijk);
  int dir = f / 2, s = f % 2;
  int c = ijk[dir] + (s ? 1 : -1);
  if (c < 0 || c >= m->n[dir]) return -1;
  long stride = dir == 0 ? 1 : dir == 1 ? m->n[0] : (long)m->n[0] * m->n[1];
  return e + (s ? stride : -stride);
}

/* testdg.cc:97: x = interpolate(|x|^2); interpolation = point evaluation at the GL nodes
 * (qkgllocalinterpolation.hh:56-74) */
void orc_interpolate_normsq(const omesh* m, double* x) {
  for (long e = 0; e < m->nelem; e++) {
    int p = m->deg[e], n = p + 1, ijk[3];
    elem_ijk(m, e, ijk);
    const deg_tab* dt = get_deg(p);
    int ne = ipow(n, m->dim);
    for (int i = 0; i < ne; i++) {
      int r = i; double s = 0;
      for (int d = 0; d < m->dim; d++) {
        int a = r % n; r /= n;
        double c = (ijk[d] + dt->nodes[a]) * m->h[d];
        s += c * c;
      }
      x[m->off[e] + i] = s;
    }
  }
}

/* ------------------------------------------------------------------------ */
/* shape function evaluation at tensor points                                */
/* ------------------------------------------------------------------------ */
/* values and reference gradients of all n_e shape functions at a point given by the 1-D
 * factor rows v[d][.] / dv[d][.]  (qkgllocalbasis.hh:91-138, x-fastest multi-index :69-78) */
static void shape_all(int dim, int n, const double* const* v, const double* const* dv,
                      double* val, double* grad /* [i][dim] */) {
  int ne = ipow(n, dim);
  for (int i = 0; i < ne; i++) {
    int a[3] = {0, 0, 0}, r = i;
    for (int d = 0; d < dim; d++) { a[d] = r % n; r /= n; }
    double pv = 1.0;
    for (int d = 0; d < dim; d++) pv *= v[d][a[d]];
    if (val) val[i] = pv;
    if (grad)
      for (int j = 0; j < dim; j++) {
        double g = dv[j][a[j]];
        for (int l = 0; l < dim; l++)
          if (l != j) g *= v[l][a[l]];
        grad[i * 3 + j] = g;
      }
  }
}

/* ------------------------------------------------------------------------ */
/* matrix-free apply: operator.hh:41-56 driving ipdgoperator.hh              */
/* ------------------------------------------------------------------------ */

/* ipdgoperator.hh:247-297 computeBulk: local[i] += (J^-T grad phi_i) . (J^-T sum_j c_j grad phi_j) w detJ */
static void mf_bulk(const omesh* m, long e, const double* x, double* local, double* gradbuf) {
  int dim = m->dim, p = m->deg[e], n = p + 1, ne = ipow(n, dim);
  int mq = p + 1; /* rule of order 2p (:251-253) -> p+1 Gauss points per direction */
  const eval_tab* et = get_eval(p, mq);
  const gauss_tab* gt = get_gauss(mq);
  const double* c = x + m->off[e];
  double detJ = 1.0;
  for (int d = 0; d < dim; d++) detJ *= m->h[d];
  int nq = ipow(mq, dim);
  for (int q = 0; q < nq; q++) {
    int qa[3] = {0, 0, 0}, r = q;
    double w = 1.0;
    const double* v[3]; const double* dv[3];
    for (int d = 0; d < dim; d++) {
      qa[d] = r % mq; r /= mq;
      w *= gt->w[qa[d]];
      v[d] = et->v + qa[d] * MAXN; dv[d] = et->d + qa[d] * MAXN;
    }
    shape_all(dim, n, v, dv, NULL, gradbuf);
    double duq[3] = {0, 0, 0};
    for (int i = 0; i < ne; i++)                           /* :278-279 */
      for (int d = 0; d < dim; d++) duq[d] += c[i] * gradbuf[i * 3 + d];
    double z = detJ * w;                                   /* :286 */
    for (int d = 0; d < dim; d++) duq[d] = duq[d] / m->h[d] * z; /* :283-287 */
    for (int i = 0; i < ne; i++) {                         /* :290-295 */
      double s = 0;
      for (int d = 0; d < dim; d++) s += gradbuf[i * 3 + d] / m->h[d] * duq[d];
      local[i] += s;
    }
  }
}

typedef struct {
  double *inV, *inG, *outV, *outG; /* shape values / reference gradients */
} facebuf;

/* Set the 1-D factor rows for a face quadrature point: normal direction pinned to the end
 * point `side`, tangential directions at Gauss points qa[]. */
static void face_rows(int dim, int dir, int side, int p, int mq, const int* qa,
                      const double** v, const double** dv) {
  const deg_tab* dt = get_deg(p);
  const eval_tab* et = get_eval(p, mq);
  int t = 0;
  for (int d = 0; d < dim; d++) {
    if (d == dir) { v[d] = dt->endv[side]; dv[d] = dt->endd[side]; }
    else { v[d] = et->v + qa[t] * MAXN; dv[d] = et->d + qa[t] * MAXN; t++; }
  }
}

/* ipdgoperator.hh:96-245: one interior face, inside = e (higher index), outside = o */
static void mf_interior_face(const omesh* m, long e, int f, long o, const double* x,
                             double* local, double* outer, facebuf* fb) {
  int dim = m->dim, dir = f / 2, side = f % 2;
  int pi = m->deg[e], po = m->deg[o], ni = pi + 1, no = po + 1;
  int nei = ipow(ni, dim), neo = ipow(no, dim);
  int maxOrder = pi > po ? pi : po;                       /* :129 */
  int pen_order = m->pdeg[e] > m->pdeg[o] ? m->pdeg[e] : m->pdeg[o];
  double penalty = m->sigma * (double)pen_order * pen_order;   /* :131 */
  int mq = maxOrder + 1;                                   /* order 2*maxOrder (:133) */
  const gauss_tab* gt = get_gauss(mq);
  double edgeLength = 1.0;                                 /* :137 */
  for (int d = 0; d < dim; d++) if (d != dir) edgeLength *= m->h[d];
  double nu = side ? 1.0 : -1.0;                           /* outer normal = nu * e_dir */
  const double* ci = x + m->off[e];
  const double* co = x + m->off[o];
  int nq = ipow(mq, dim - 1);
  for (int q = 0; q < nq; q++) {
    int qa[2] = {0, 0}, r = q; double w = 1.0;
    for (int t = 0; t < dim - 1; t++) { qa[t] = r % mq; r /= mq; w *= gt->w[qa[t]]; }
    const double *v[3], *dv[3];
    face_rows(dim, dir, side, pi, mq, qa, v, dv);
    shape_all(dim, ni, v, dv, fb->inV, fb->inG);           /* :154-155 */
    face_rows(dim, dir, 1 - side, po, mq, qa, v, dv);
    shape_all(dim, no, v, dv, fb->outV, fb->outG);         /* :156-157 */
    double inDn = 0, outDn = 0, inU = 0, outU = 0;
    /* only the normal component of J^-T grad survives the product with the normal */
    for (int i = 0; i < nei; i++) { inDn += ci[i] * fb->inG[i * 3 + dir]; inU += ci[i] * fb->inV[i]; }   /* :161-164 */
    for (int i = 0; i < neo; i++) { outDn += co[i] * fb->outG[i * 3 + dir]; outU += co[i] * fb->outV[i]; } /* :165-168 */
    inDn /= m->h[dir]; outDn /= m->h[dir];                 /* :176-177 */
    double weight = edgeLength * w;                        /* :173 */
    /* 0. -{du/dn}[phi] (:181-192) */
    {
      double avg = (inDn + outDn) * nu * 0.5 * weight;
      for (int i = 0; i < nei; i++) local[i] -= avg * fb->inV[i];
      for (int i = 0; i < neo; i++) outer[i] += avg * fb->outV[i];
    }
    /* 1. -{dphi/dn}[u] (:194-215) */
    {
      double jump = (inU - outU) * weight;
      for (int i = 0; i < nei; i++) local[i] -= jump * (0.5 * fb->inG[i * 3 + dir] / m->h[dir] * nu);
      for (int i = 0; i < neo; i++) outer[i] -= jump * (0.5 * fb->outG[i * 3 + dir] / m->h[dir] * nu);
    }
    /* 2. + sigma/|e| [phi][u] (:217-231) */
    {
      double jump = (inU - outU) * (weight * penalty / edgeLength);
      for (int i = 0; i < nei; i++) local[i] += jump * fb->inV[i];
      for (int i = 0; i < neo; i++) outer[i] -= jump * fb->outV[i];
    }
  }
}

/* ipdgoperator.hh:302-390 computeDirichletBoundaryEdge */
static void mf_boundary_face(const omesh* m, long e, int f, const double* x, double* local, facebuf* fb) {
  int dim = m->dim, dir = f / 2, side = f % 2;
  int p = m->deg[e], n = p + 1, ne = ipow(n, dim);
  double penalty = m->sigma * (double)m->pdeg[e] * m->pdeg[e];  /* :310 */
  int mq = p + 1;                                               /* :315-318 */
  const gauss_tab* gt = get_gauss(mq);
  double edgeLength = 1.0;
  for (int d = 0; d < dim; d++) if (d != dir) edgeLength *= m->h[d];
  double nu = side ? 1.0 : -1.0;
  const double* ci = x + m->off[e];
  int nq = ipow(mq, dim - 1);
  for (int q = 0; q < nq; q++) {
    int qa[2] = {0, 0}, r = q; double w = 1.0;
    for (int t = 0; t < dim - 1; t++) { qa[t] = r % mq; r /= mq; w *= gt->w[qa[t]]; }
    const double *v[3], *dv[3];
    face_rows(dim, dir, side, p, mq, qa, v, dv);
    shape_all(dim, n, v, dv, fb->inV, fb->inG);
    double inDn = 0, inU = 0;
    for (int i = 0; i < ne; i++) { inDn += ci[i] * fb->inG[i * 3 + dir]; inU += ci[i] * fb->inV[i]; }
    inDn /= m->h[dir];
    double weight = edgeLength * w;
    double avg = inDn * nu * weight;                             /* :355-357 */
    for (int i = 0; i < ne; i++) local[i] -= avg * fb->inV[i];
    double jump = inU * weight;                                  /* :365-374 */
    for (int i = 0; i < ne; i++) local[i] -= jump * (fb->inG[i * 3 + dir] / m->h[dir] * nu);
    double jp = inU * (weight * penalty / edgeLength);           /* :379-386 */
    for (int i = 0; i < ne; i++) local[i] += jp * fb->inV[i];
  }
}

static int max_ne(const omesh* m) {
  int mx = 0;
  for (long e = 0; e < m->nelem; e++) if (m->deg[e] > mx) mx = m->deg[e];
  return ipow(mx + 1, m->dim);
}

static void mf_element(const omesh* m, long e, const double* x, double* y, double factor,
                       double* local, double* outer, double* gradbuf, facebuf* fb) {
  int ne = (int)(m->off[e + 1] - m->off[e]);
  for (int i = 0; i < ne; i++) local[i] = 0;               /* bind(): ipdgoperator.hh:48-55 */
  mf_bulk(m, e, x, local, gradbuf);                        /* compute(): :57-60 */
  for (int f = 0; f < 2 * m->dim; f++) {                   /* computeFace(): :96 */
    long o = neighbour(m, e, f);
    if (o < 0) {
      if (m->dirichlet) mf_boundary_face(m, e, f, x, local, fb); /* :97-105 */
      continue;
    }
    if (e < o) continue;                                   /* :108 */
    int neo = (int)(m->off[o + 1] - m->off[o]);
    for (int i = 0; i < neo; i++) outer[i] = 0;            /* :114-116 */
    mf_interior_face(m, e, f, o, x, local, outer, fb);
    for (int i = 0; i < neo; i++) y[m->off[o] + i] += factor * outer[i]; /* :234-243 */
  }
  if (factor != 0.0)                                       /* write(): :62-76 */
    for (int i = 0; i < ne; i++) y[m->off[e] + i] += factor * local[i];
}

void orc_apply_mf(const omesh* m, const double* x, double* y, double factor, int threads) {
  long N = orc_mesh_dimension(m);
  int mne = max_ne(m);
  memset(y, 0, sizeof(double) * N);                        /* operator.hh:42 */
  if (threads <= 1) {
    double* buf = (double*)malloc(sizeof(double) * mne * 10);
    facebuf fb = {buf + 2 * mne, buf + 3 * mne, buf + 6 * mne, buf + 7 * mne};
    double* gradbuf = (double*)malloc(sizeof(double) * mne * 3);
    for (long e = 0; e < m->nelem; e++)                    /* operator.hh:49-55 */
      mf_element(m, e, x, y, factor, buf, buf + mne, gradbuf, &fb);
    free(buf); free(gradbuf);
    return;
  }
  /* Threaded variant for the timing baseline: the reference is single-threaded and its
   * scatter into the neighbour's rows is "not thread-safe" (ipdgoperator.hh:233).  An
   * element writes its own rows and those of its lower-index face neighbours, so elements
   * whose indices have equal parity in every direction never collide: 2^dim colours. */
  for (int p = 0; p <= MAXP; p++) { /* warm the caches outside the parallel region */
    int used = 0;
    for (long e = 0; e < m->nelem && !used; e++) used = (m->deg[e] == p);
    if (!used) continue;
    get_deg(p);
    for (int q = 0; q <= MAXP; q++) {
      int u2 = 0;
      for (long e = 0; e < m->nelem && !u2; e++) u2 = (m->deg[e] == q);
      if (u2 && q >= p) { get_eval(p, q + 1); get_eval(q, q + 1); }
    }
  }
#ifdef _OPENMP
  int nthreads = threads;
#endif
  for (int colour = 0; colour < (1 << m->dim); colour++) {
#pragma omp parallel num_threads(nthreads)
    {
      double* buf = (double*)malloc(sizeof(double) * mne * 10);
      facebuf fb = {buf + 2 * mne, buf + 3 * mne, buf + 6 * mne, buf + 7 * mne};
      double* gradbuf = (double*)malloc(sizeof(double) * mne * 3);
#pragma omp for schedule(static)
      for (long e = 0; e < m->nelem; e++) {
        int ijk[3]; elem_ijk(m, e, ijk);
        int c = (ijk[0] & 1) | ((ijk[1] & 1) << 1) | ((ijk[2] & 1) << 2);
        if (c != colour) continue;
        mf_element(m, e, x, y, factor, buf, buf + mne, gradbuf, &fb);
      }
      free(buf); free(gradbuf);
    }
  }
}

/* ------------------------------------------------------------------------ */
/* assembled matrix in DynamicBCRSMatrix layout                              */
/* ------------------------------------------------------------------------ */
struct obcrs {
  long nrows;        /* block rows */
  long* rowptr;      /* nrows+1 */
  int* col;          /* block column indices, ascending per row */
  long* boff;        /* nblocks+1 : offset of block into val (dynamicbcrs.hh:190-199) */
  int* brows;        /* rowMap_ (dynamicbcrs.hh:207) */
  int* bcols;        /* colMap_ */
  long* voff;        /* vector offsets, nrows+1 */
  double* val;       /* one contiguous array (dynamicbcrs.hh:211) */
};

static obcrs* bcrs_alloc_pattern(const omesh* m, const int* bsize) {
  /* pattern = element + face neighbours (matrices.hh:42-44 assembleSkeletonPattern);
   * ISTL BCRS rows keep column indices ascending */
  obcrs* A = (obcrs*)calloc(1, sizeof(obcrs));
  long ne = m->nelem;
  A->nrows = ne;
  A->rowptr = (long*)malloc(sizeof(long) * (ne + 1));
  A->brows = (int*)malloc(sizeof(int) * ne);
  A->bcols = (int*)malloc(sizeof(int) * ne);
  A->voff = (long*)malloc(sizeof(long) * (ne + 1));
  A->voff[0] = 0;
  for (long e = 0; e < ne; e++) { A->brows[e] = A->bcols[e] = bsize[e]; A->voff[e + 1] = A->voff[e] + bsize[e]; }
  long nb = 0;
  A->rowptr[0] = 0;
  for (long e = 0; e < ne; e++) {
    int c = 1;
    for (int f = 0; f < 2 * m->dim; f++) {
      long o = neighbour(m, e, f);
      if (o >= 0) {
        /* a periodic-free structured mesh with n=2 in a direction has distinct -/+ neighbours
         * only when n>2; with n==1 there is none, with n==2 exactly one */
        c++;
      }
    }
    nb += c; A->rowptr[e + 1] = nb;
  }
  A->col = (int*)malloc(sizeof(int) * nb);
  A->boff = (long*)malloc(sizeof(long) * (nb + 1));
  long k = 0, off = 0;
  for (long e = 0; e < ne; e++) {
    long cols[7]; int c = 0;
    cols[c++] = e;
    for (int f = 0; f < 2 * m->dim; f++) { long o = neighbour(m, e, f); if (o >= 0) cols[c++] = o; }
    for (int a = 1; a < c; a++) { long v = cols[a]; int b = a - 1; while (b >= 0 && cols[b] > v) { cols[b + 1] = cols[b]; b--; } cols[b + 1] = v; }
    for (int a = 0; a < c; a++) {
      A->col[k] = (int)cols[a];
      A->boff[k] = off;
      off += (long)bsize[e] * bsize[cols[a]];          /* dynamicbcrs.hh:178-187 calculateSize */
      k++;
    }
  }
  A->boff[nb] = off;
  A->val = (double*)calloc((size_t)off, sizeof(double));  /* matrices.hh:49 matrix = 0 */
  return A;
}

void orc_bcrs_destroy(obcrs* A) {
  if (!A) return;
  free(A->rowptr); free(A->col); free(A->boff); free(A->brows); free(A->bcols); free(A->voff); free(A->val); free(A);
}
long orc_bcrs_nblocks(const obcrs* A) { return A->rowptr[A->nrows]; }
long orc_bcrs_nentries(const obcrs* A) { return A->boff[A->rowptr[A->nrows]]; }
void orc_bcrs_export(const obcrs* A, long* rowptr, int* col, long* boff, double* val) {
  long nb = A->rowptr[A->nrows];
  if (rowptr) memcpy(rowptr, A->rowptr, sizeof(long) * (A->nrows + 1));
  if (col) memcpy(col, A->col, sizeof(int) * nb);
  if (boff) memcpy(boff, A->boff, sizeof(long) * (nb + 1));
  if (val) memcpy(val, A->val, sizeof(double) * A->boff[nb]);
}
static double* bcrs_block(const obcrs* A, long i, long j) {
  for (long k = A->rowptr[i]; k < A->rowptr[i + 1]; k++)
    if (A->col[k] == j) return A->val + A->boff[k];
  return NULL;
}

/* dune-fufem LaplaceAssembler (un-vendored; call sites test/testobjects.hh:68,
 * ipdgblockjacobi.hh:62-66): local[i][j] += (J^-T grad phi_i).(J^-T grad phi_j) w detJ */
static void asm_bulk(const omesh* m, long e, double* blk, int ld, double* gradbuf) {
  int dim = m->dim, p = m->deg[e], n = p + 1, ne = ipow(n, dim);
  int mq = p + 1;
  const eval_tab* et = get_eval(p, mq);
  const gauss_tab* gt = get_gauss(mq);
  double detJ = 1.0;
  for (int d = 0; d < dim; d++) detJ *= m->h[d];
  int nq = ipow(mq, dim);
  for (int q = 0; q < nq; q++) {
    int r = q; double w = 1.0; const double *v[3], *dv[3];
    for (int d = 0; d < dim; d++) { int a = r % mq; r /= mq; w *= gt->w[a]; v[d] = et->v + a * MAXN; dv[d] = et->d + a * MAXN; }
    shape_all(dim, n, v, dv, NULL, gradbuf);
    double z = w * detJ;
    for (int i = 0; i < ne; i++)
      for (int j = 0; j < ne; j++) {
        double s = 0;
        for (int d = 0; d < dim; d++) s += gradbuf[i * 3 + d] * gradbuf[j * 3 + d] / (m->h[d] * m->h[d]);
        blk[i * ld + j] += s * z;
      }
  }
}

/* variableipdg.hh:249-365 assembleBlockwise (SIPG: dgType_ = -1, sigma1_ = 0).
 * M11 -> Aii, M12 -> Aio, M21 -> Aoi, M22 -> Aoo; any of the four may be NULL. */
static void asm_interior_face(const omesh* m, long e, int f, long o,
                              double* Aii, double* Aio, double* Aoi, double* Aoo, facebuf* fb) {
  int dim = m->dim, dir = f / 2, side = f % 2;
  int pi = m->deg[e], po = m->deg[o], ni = pi + 1, no = po + 1;
  int nei = ipow(ni, dim), neo = ipow(no, dim);
  int maxOrder = pi > po ? pi : po;                         /* :252-254, :271 */
  int pen_order = m->pdeg[e] > m->pdeg[o] ? m->pdeg[e] : m->pdeg[o];
  double penalty = m->sigma * (double)pen_order * pen_order;    /* :255 */
  int mq = maxOrder + 1;
  const gauss_tab* gt = get_gauss(mq);
  double edgeLength = 1.0;
  for (int d = 0; d < dim; d++) if (d != dir) edgeLength *= m->h[d];
  double nu = side ? 1.0 : -1.0;
  const double dgType = -1.0;
  int nq = ipow(mq, dim - 1);
  for (int q = 0; q < nq; q++) {
    int qa[2] = {0, 0}, r = q; double w = 1.0;
    for (int t = 0; t < dim - 1; t++) { qa[t] = r % mq; r /= mq; w *= gt->w[qa[t]]; }
    const double *v[3], *dv[3];
    face_rows(dim, dir, side, pi, mq, qa, v, dv);
    shape_all(dim, ni, v, dv, fb->inV, fb->inG);
    face_rows(dim, dir, 1 - side, po, mq, qa, v, dv);
    shape_all(dim, no, v, dv, fb->outV, fb->outG);
    double z = w * edgeLength;                              /* :318 */
    double pz = penalty * z / edgeLength;
#define GN_IN(i) (fb->inG[(i) * 3 + dir] / m->h[dir] * nu)
#define GN_OUT(i) (fb->outG[(i) * 3 + dir] / m->h[dir] * nu)
    for (int i = 0; i < nei; i++) {
      if (Aii) for (int j = 0; j < nei; j++)                /* M11 :327-331 */
        Aii[i * nei + j] += -0.5 * z * fb->inV[i] * GN_IN(j) + 0.5 * dgType * z * fb->inV[j] * GN_IN(i) + pz * fb->inV[i] * fb->inV[j];
      if (Aio) for (int j = 0; j < neo; j++)                /* M12 :336-340 */
        Aio[i * neo + j] += -0.5 * z * fb->inV[i] * GN_OUT(j) - 0.5 * dgType * z * fb->outV[j] * GN_IN(i) - pz * fb->inV[i] * fb->outV[j];
    }
    for (int i = 0; i < neo; i++) {
      if (Aoi) for (int j = 0; j < nei; j++)                /* M21 :348-352 */
        Aoi[i * nei + j] += 0.5 * z * fb->outV[i] * GN_IN(j) + 0.5 * dgType * z * fb->inV[j] * GN_OUT(i) - pz * fb->outV[i] * fb->inV[j];
      if (Aoo) for (int j = 0; j < neo; j++)                /* M22 :357-361 */
        Aoo[i * neo + j] += 0.5 * z * fb->outV[i] * GN_OUT(j) - 0.5 * dgType * z * fb->outV[j] * GN_OUT(i) + pz * fb->outV[i] * fb->outV[j];
    }
  }
}

/* variableipdg.hh:101-184 boundary face (Dirichlet only) */
static void asm_boundary_face(const omesh* m, long e, int f, double* A, facebuf* fb) {
  if (!m->dirichlet) return;                                /* :109-110 */
  int dim = m->dim, dir = f / 2, side = f % 2;
  int p = m->deg[e], n = p + 1, ne = ipow(n, dim);
  double penalty = m->sigma * (double)m->pdeg[e] * m->pdeg[e];  /* :115 */
  int mq = p + 1;
  const gauss_tab* gt = get_gauss(mq);
  double edgeLength = 1.0;
  for (int d = 0; d < dim; d++) if (d != dir) edgeLength *= m->h[d];
  double nu = side ? 1.0 : -1.0;
  const double dgType = -1.0;
  int nq = ipow(mq, dim - 1);
  for (int q = 0; q < nq; q++) {
    int qa[2] = {0, 0}, r = q; double w = 1.0;
    for (int t = 0; t < dim - 1; t++) { qa[t] = r % mq; r /= mq; w *= gt->w[qa[t]]; }
    const double *v[3], *dv[3];
    face_rows(dim, dir, side, p, mq, qa, v, dv);
    shape_all(dim, n, v, dv, fb->inV, fb->inG);
    double z = w * edgeLength;
    for (int i = 0; i < ne; i++)
      for (int j = 0; j < ne; j++)                           /* :176-178 */
        A[i * ne + j] += -z * fb->inV[i] * GN_IN(j) + dgType * z * fb->inV[j] * GN_IN(i) + penalty * z / edgeLength * fb->inV[i] * fb->inV[j];
  }
}

obcrs* orc_assemble(const omesh* m) {
  long ne = m->nelem;
  int* bs = (int*)malloc(sizeof(int) * ne);
  for (long e = 0; e < ne; e++) bs[e] = (int)(m->off[e + 1] - m->off[e]);
  obcrs* A = bcrs_alloc_pattern(m, bs);
  free(bs);
  int mne = max_ne(m);
#pragma omp parallel
  {
    double* buf = (double*)malloc(sizeof(double) * mne * 8);
    facebuf fb = {buf, buf + mne, buf + 4 * mne, buf + 5 * mne};
    double* gradbuf = (double*)malloc(sizeof(double) * mne * 3);
    /* Row-wise ("pull") accumulation so rows can be built in parallel: block row e receives
     * the bulk term, and for every face the blocks that the face assembler produces for test
     * functions living on e.  Inside = the higher-index element as in ipdgoperator.hh:108
     * (the SIPG form is orientation independent). */
#pragma omp for schedule(dynamic, 16)
    for (long e = 0; e < ne; e++) {
      double* Aee = bcrs_block(A, e, e);
      int nee = A->brows[e];
      asm_bulk(m, e, Aee, nee, gradbuf);                     /* matrices.hh:78-83 */
      for (int f = 0; f < 2 * m->dim; f++) {
        long o = neighbour(m, e, f);
        if (o < 0) { asm_boundary_face(m, e, f, Aee, &fb); continue; }  /* matrices.hh:68-76 */
        if (e > o) asm_interior_face(m, e, f, o, Aee, bcrs_block(A, e, o), NULL, NULL, &fb);
        else asm_interior_face(m, o, f ^ 1, e, NULL, NULL, bcrs_block(A, e, o), Aee, &fb);
      }
    }
    free(buf); free(gradbuf);
  }
  return A;
}

/* y = A x: BCRSMatrix::mv over MatrixWindow::umv (matrixwindow.hh:196-209) */
void orc_bcrs_mv(const obcrs* A, const double* x, double* y, int threads) {
#pragma omp parallel for schedule(static) if (threads > 1) num_threads(threads > 1 ? threads : 1)
  for (long i = 0; i < A->nrows; i++) {
    int r = A->brows[i];
    double* yi = y + A->voff[i];
    for (int a = 0; a < r; a++) yi[a] = 0;
    for (long k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) {
      int j = A->col[k], c = A->bcols[j];
      const double* B = A->val + A->boff[k];
      const double* xj = x + A->voff[j];
      for (int a = 0; a < r; a++) {
        double s = 0;
        for (int b = 0; b < c; b++) s += B[a * c + b] * xj[b];
        yi[a] += s;
      }
    }
  }
}
/* y -= A x (matrixwindow.hh:222-234) */
void orc_bcrs_mmv(const obcrs* A, const double* x, double* y) {
  for (long i = 0; i < A->nrows; i++) {
    int r = A->brows[i];
    double* yi = y + A->voff[i];
    for (long k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) {
      int j = A->col[k], c = A->bcols[j];
      const double* B = A->val + A->boff[k];
      const double* xj = x + A->voff[j];
      for (int a = 0; a < r; a++)
        for (int b = 0; b < c; b++) yi[a] -= B[a * c + b] * xj[b];
    }
  }
}
double orc_bcrs_frobenius_diff(const obcrs* A, const obcrs* B) {
  long n = orc_bcrs_nentries(A);
  if (n != orc_bcrs_nentries(B)) return INFINITY;
  double s = 0;
  for (long i = 0; i < n; i++) { double d = A->val[i] - B->val[i]; s += d * d; }
  return sqrt(s);
}
void orc_bcrs_diag_block(const obcrs* A, long e, double* out) {
  const double* B = bcrs_block(A, e, e);
  memcpy(out, B, sizeof(double) * A->brows[e] * A->bcols[e]);
}

/* ------------------------------------------------------------------------ */
/* smoothers                                                                 */
/* ------------------------------------------------------------------------ */
/* Imp::GSCore (dynamicblockgs.hh:17-40): one forward scalar GS sweep from zero */
static void gs_core(const double* M, int n, const double* b, double* x) {
  for (int i = 0; i < n; i++) x[i] = 0;
  for (int i = 0; i < n; i++) {
    double mii = M[i * n + i];
    if (fabs(mii) == 0.) continue;
    double xi = b[i];
    for (int j = 0; j < n; j++) if (j != i) xi -= M[i * n + j] * x[j];
    x[i] = xi / mii;
  }
}

/* DynamicBlockGS::iterate (dynamicblockgs.hh:94-126) */
void orc_blockgs_iterate(const obcrs* A, const double* b, double* x) {
  int mx = 0;
  for (long i = 0; i < A->nrows; i++) if (A->brows[i] > mx) mx = A->brows[i];
  double* ri = (double*)malloc(sizeof(double) * mx * 2);
  double* corr = ri + mx;
  for (long i = 0; i < A->nrows; i++) {
    int r = A->brows[i];
    memcpy(ri, b + A->voff[i], sizeof(double) * r);          /* r = copy(b) :98 */
    const double* diag = NULL;
    for (long k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) { /* :108-111, whole row incl. diagonal */
      int j = A->col[k], c = A->bcols[j];
      const double* B = A->val + A->boff[k];
      const double* xj = x + A->voff[j];
      if (j == i) diag = B;
      for (int a = 0; a < r; a++)
        for (int bb = 0; bb < c; bb++) ri[a] -= B[a * c + bb] * xj[bb];
    }
    gs_core(diag, r, ri, corr);                              /* :121 */
    for (int a = 0; a < r; a++) x[A->voff[i] + a] += corr[a]; /* :122-124 */
  }
  free(ri);
}

/* L1Smoother::preprocess (iterationsteps/l1smoother.hh:31-57): for every ghost block index g (list order, duplicates
 * count again) and every block row r != g in the pattern of row g, add to reg_r[j] the l1 norm of row j of block A[r][g]. */
void orc_l1_regularization(const obcrs* A, const long* ghosts, long nghost, double* reg) {
  long N = A->voff[A->nrows];
  for (long i = 0; i < N; i++) reg[i] = 0.0;                                  /* :33 */
  for (long q = 0; q < nghost; q++) {
    long g = ghosts[q];
    for (long k = A->rowptr[g]; k < A->rowptr[g + 1]; k++) {                  /* :41 */
      long row = A->col[k];
      if (row == g) continue;                                                 /* :45-46 */
      const double* B = NULL; int c = 0;
      for (long kk = A->rowptr[row]; kk < A->rowptr[row + 1]; kk++)           /* m[row][ghostIdx] :47 */
        if (A->col[kk] == g) { B = A->val + A->boff[kk]; c = A->bcols[g]; }
      if (!B) continue;
      int r = A->brows[row];
      for (int j = 0; j < r; j++)                                             /* :49-54 */
        for (int e = 0; e < c; e++) reg[A->voff[row] + j] += fabs(B[j * c + e]);
    }
  }
}

/* L1Smoother::gs (l1smoother.hh:127-145): forward scalar GS sweep from zero on (M + diag(d)), tolerance 0 */
static void gs_l1(const double* M, int n, const double* b, const double* d, double* x) {
  for (int i = 0; i < n; i++) x[i] = 0;
  for (int i = 0; i < n; i++) {
    double mii = M[i * n + i];
    if (fabs(mii) <= 0.0) continue;                                           /* :136-137 */
    double xi = b[i];
    for (int j = 0; j < n; j++) if (j != i) xi -= M[i * n + j] * x[j];
    x[i] = xi / (mii + d[i]);                                                 /* :144 */
  }
}

/* L1Smoother::iterate (l1smoother.hh:63-113): block rows in ascending order (ghost rows included, :71-73) */
void orc_l1_iterate(const obcrs* A, const double* reg, const double* b, double* x) {
  int mx = 0;
  for (long i = 0; i < A->nrows; i++) if (A->brows[i] > mx) mx = A->brows[i];
  double* ri = (double*)malloc(sizeof(double) * mx * 2);
  double* corr = ri + mx;
  for (long i = 0; i < A->nrows; i++) {
    int r = A->brows[i];
    memcpy(ri, b + A->voff[i], sizeof(double) * r);                           /* r = b :67 */
    const double* diag = NULL;
    for (long k = A->rowptr[i]; k < A->rowptr[i + 1]; k++) {                  /* :91-94 */
      int j = A->col[k], c = A->bcols[j];
      const double* B = A->val + A->boff[k];
      const double* xj = x + A->voff[j];
      if (j == i) diag = B;
      for (int a = 0; a < r; a++)
        for (int bb = 0; bb < c; bb++) ri[a] -= B[a * c + bb] * xj[bb];
    }
    gs_l1(diag, r, ri, reg + A->voff[i], corr);                               /* :105 */
    for (int a = 0; a < r; a++) x[A->voff[i] + a] += corr[a];                 /* :107-109 */
  }
  free(ri);
}

/* ipdgblockjacobi.hh:58-152: the diagonal block as the matrix-free block Jacobi builds it */
void orc_diag_block_mf(const omesh* m, long e, double* out) {
  int dim = m->dim, p = m->deg[e], n = p + 1, ne = ipow(n, dim);
  int mne = max_ne(m);
  double* buf = (double*)malloc(sizeof(double) * mne * 8);
  facebuf fb = {buf, buf + mne, buf + 4 * mne, buf + 5 * mne};
  double* gradbuf = (double*)malloc(sizeof(double) * mne * 3);
  memset(out, 0, sizeof(double) * ne * ne);
  asm_bulk(m, e, out, ne, gradbuf);                          /* :62-66 */
  for (int f = 0; f < 2 * dim; f++) {                        /* :68 */
    long o = neighbour(m, e, f);
    double avg_factor = 0.5;
    if (o < 0) { if (!m->dirichlet) continue; avg_factor = 1.0; }   /* :69-76 */
    int order = p, pen_order = m->pdeg[e];
    if (o >= 0) {                                            /* :80-85 */
      if (m->deg[o] > order) order = m->deg[o];
      if (m->pdeg[o] > pen_order) pen_order = m->pdeg[o];
    }
    double penalty = m->sigma * (double)pen_order * pen_order;   /* :86 */
    int dir = f / 2, side = f % 2;
    int mq = order + 1;                                      /* :100-103 */
    const gauss_tab* gt = get_gauss(mq);
    double edgeLength = 1.0;
    for (int d = 0; d < dim; d++) if (d != dir) edgeLength *= m->h[d];
    double nu = side ? 1.0 : -1.0;
    int nq = ipow(mq, dim - 1);
    for (int q = 0; q < nq; q++) {
      int qa[2] = {0, 0}, r = q; double w = 1.0;
      for (int t = 0; t < dim - 1; t++) { qa[t] = r % mq; r /= mq; w *= gt->w[qa[t]]; }
      const double *v[3], *dv[3];
      face_rows(dim, dir, side, p, mq, qa, v, dv);
      shape_all(dim, n, v, dv, fb.inV, fb.inG);
      double z = w * edgeLength;
      facebuf* fbp = &fb;
#define GN2(i) (fbp->inG[(i) * 3 + dir] / m->h[dir] * nu)
      for (int i = 0; i < ne; i++)
        for (int j = 0; j < ne; j++)                         /* :141-146 */
          out[i * ne + j] += -avg_factor * z * fb.inV[i] * GN2(j) - avg_factor * z * fb.inV[j] * GN2(i) + penalty * z / edgeLength * fb.inV[i] * fb.inV[j];
    }
  }
  free(buf); free(gradbuf);
}

/* dense Cholesky solve (the "exact local solver" the north star asks for) */
static int chol_solve(double* A, int n, double* b) {
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
    if (d <= 0) return 1;
    d = sqrt(d); A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= A[i * n + k] * b[k]; b[i] = s / A[i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s -= A[k * n + i] * b[k]; b[i] = s / A[i * n + i]; }
  return 0;
}

/* Operator::apply with an IPDGBlockJacobi local operator: c = factor * sum_e P_e^T solve(D_e, P_e r)
 * (operator.hh:41-56, ipdgblockjacobi.hh:154-178) */
void orc_blockjacobi_apply(const omesh* m, const double* r, double* c, double factor, int local_solver) {
  int mne = max_ne(m);
#pragma omp parallel
  {
    double* D = (double*)malloc(sizeof(double) * mne * mne);
    double* v = (double*)malloc(sizeof(double) * mne);
#pragma omp for schedule(dynamic, 8)
    for (long e = 0; e < m->nelem; e++) {
      int ne = (int)(m->off[e + 1] - m->off[e]);
      orc_diag_block_mf(m, e, D);
      if (local_solver == 0) {
        memcpy(v, r + m->off[e], sizeof(double) * ne);
        chol_solve(D, ne, v);
      } else {
        gs_core(D, ne, r + m->off[e], v);                     /* testdgblockjacobi.cc:63-76 */
      }
      for (int i = 0; i < ne; i++) c[m->off[e] + i] = factor * v[i];
    }
    free(D); free(v);
  }
}

/* ------------------------------------------------------------------------ */
/* p-transfer                                                                */
/* ------------------------------------------------------------------------ */
/* TransferMatrixCache::makeTransferMatrix (dynamicordertransfer.hh:48-73):
 * T[i][j] = phi^coarse_j(x^fine_i) */
void orc_transfer_matrix(int dim, int pc, int pf, double* T) {
  const deg_tab* dc = get_deg(pc);
  const deg_tab* df = get_deg(pf);
  int nc = pc + 1, nf = pf + 1;
  int nec = ipow(nc, dim), nef = ipow(nf, dim);
  for (int i = 0; i < nef; i++)
    for (int j = 0; j < nec; j++) {
      int ri = i, rj = j; double v = 1.0;
      for (int d = 0; d < dim; d++) {
        int a = ri % nf, b = rj % nc; ri /= nf; rj /= nc;
        v *= orc_lagrange(pc, dc->nodes, b, df->nodes[a]);
      }
      T[i * nec + j] = v;
    }
}

/* DGOrderTransfer::setup (ordertransfer.hh:45-88): blocks above maxOrder are cut to maxOrder,
 * the others keep their size (identity).  Penalty degrees are inherited (Galerkin product). */
omesh* orc_mesh_coarsen(const omesh* m, int max_order) {
  int* deg = (int*)malloc(sizeof(int) * m->nelem);
  for (long e = 0; e < m->nelem; e++) deg[e] = m->deg[e] < max_order ? m->deg[e] : max_order;
  omesh* c = orc_mesh_create(m->dim, m->n, m->L, deg, m->pdeg, m->sigma, m->dirichlet);
  free(deg);
  return c;
}

/* restrict: coarse = T^T fine (ordertransfer.hh:91-102, arithmetic.hh transposedMatrixVectorProduct) */
void orc_restrict(const omesh* fine, const omesh* coarse, const double* xf, double* xc) {
  double* T = (double*)malloc(sizeof(double) * max_ne(fine) * max_ne(coarse));
  for (long e = 0; e < fine->nelem; e++) {
    int nf = (int)(fine->off[e + 1] - fine->off[e]), nc = (int)(coarse->off[e + 1] - coarse->off[e]);
    const double* f = xf + fine->off[e]; double* c = xc + coarse->off[e];
    if (nf == nc) { memcpy(c, f, sizeof(double) * nf); continue; }    /* identity block :73-77 */
    orc_transfer_matrix(fine->dim, coarse->deg[e], fine->deg[e], T);
    for (int j = 0; j < nc; j++) { double s = 0; for (int i = 0; i < nf; i++) s += T[i * nc + j] * f[i]; c[j] = s; }
  }
  free(T);
}
/* prolong: fine = T coarse (ordertransfer.hh:106-119) */
void orc_prolong(const omesh* fine, const omesh* coarse, const double* xc, double* xf) {
  double* T = (double*)malloc(sizeof(double) * max_ne(fine) * max_ne(coarse));
  for (long e = 0; e < fine->nelem; e++) {
    int nf = (int)(fine->off[e + 1] - fine->off[e]), nc = (int)(coarse->off[e + 1] - coarse->off[e]);
    double* f = xf + fine->off[e]; const double* c = xc + coarse->off[e];
    if (nf == nc) { memcpy(f, c, sizeof(double) * nf); continue; }
    orc_transfer_matrix(fine->dim, coarse->deg[e], fine->deg[e], T);
    for (int i = 0; i < nf; i++) { double s = 0; for (int j = 0; j < nc; j++) s += T[i * nc + j] * c[j]; f[i] = s; }
  }
  free(T);
}

/* galerkinRestrict (ordertransfer.hh:124-144; arithmetic.hh:93-118 addTransformedMatrix):
 * A_c[i][j] = T_i^T A_f[i][j] T_j */
obcrs* orc_galerkin_restrict(const omesh* fine, const omesh* coarse, const obcrs* Af) {
  long ne = fine->nelem;
  int* bs = (int*)malloc(sizeof(int) * ne);
  for (long e = 0; e < ne; e++) bs[e] = (int)(coarse->off[e + 1] - coarse->off[e]);
  obcrs* Ac = bcrs_alloc_pattern(coarse, bs);
  free(bs);
  int mf = max_ne(fine), mc = max_ne(coarse);
  double* Ti = (double*)malloc(sizeof(double) * mf * mc);
  double* Tj = (double*)malloc(sizeof(double) * mf * mc);
  double* tmp = (double*)malloc(sizeof(double) * mf * mc);
  for (long i = 0; i < ne; i++) {
    int nfi = Af->brows[i], nci = Ac->brows[i];
    if (nfi != nci) orc_transfer_matrix(fine->dim, coarse->deg[i], fine->deg[i], Ti);
    else { memset(Ti, 0, sizeof(double) * nfi * nci); for (int a = 0; a < nfi; a++) Ti[a * nci + a] = 1.0; }
    for (long k = Af->rowptr[i]; k < Af->rowptr[i + 1]; k++) {
      int j = Af->col[k]; int nfj = Af->bcols[j], ncj = Ac->bcols[j];
      if (nfj != ncj) orc_transfer_matrix(fine->dim, coarse->deg[j], fine->deg[j], Tj);
      else { memset(Tj, 0, sizeof(double) * nfj * ncj); for (int a = 0; a < nfj; a++) Tj[a * ncj + a] = 1.0; }
      const double* B = Af->val + Af->boff[k];
      double* C = Ac->val + Ac->boff[k];
      /* tmp (nfi x ncj) = B Tj ; C = Ti^T tmp */
      for (int a = 0; a < nfi; a++)
        for (int b = 0; b < ncj; b++) { double s = 0; for (int c = 0; c < nfj; c++) s += B[a * nfj + c] * Tj[c * ncj + b]; tmp[a * ncj + b] = s; }
      for (int a = 0; a < nci; a++)
        for (int b = 0; b < ncj; b++) { double s = 0; for (int c = 0; c < nfi; c++) s += Ti[c * nci + a] * tmp[c * ncj + b]; C[a * ncj + b] += s; }
    }
  }
  free(Ti); free(Tj); free(tmp);
  return Ac;
}

/* ------------------------------------------------------------------------ */
/* multigrid V-cycle (mg/multigrid_impl.hh:16-117)                           */
/* ------------------------------------------------------------------------ */
typedef struct {
  int nlev; omesh* const* lev; obcrs* const* mats; int smoother; double damping;
  int pre, post, coarse_its;
  double **x, **r; double *tmp1, *tmp2;
} vstate;

static void lvl_apply(const vstate* s, int l, const double* x, double* y) {
  if (s->mats) orc_bcrs_mv(s->mats[l], x, y, 1);            /* operatorFromMatrix, multigrid.hh:137-155 */
  else orc_apply_mf(s->lev[l], x, y, 1.0, orc_max_threads());
}
/* Smoother<V>(c, r): c is a correction computed from a zero start.
 * smoother 0: smootherFromIterationStep2 around DynamicBlockGS (multigrid.hh:96-107; x enters
 * as the zeroed tmp1 of applySmoother, multigrid_impl.hh:73).  smoother 1: damped exact block Jacobi. */
static void lvl_smooth(const vstate* s, int l, double* c, const double* r) {
  long N = orc_mesh_dimension(s->lev[l]);
  if (s->smoother == 0) {
    orc_blockgs_iterate(s->mats[l], r, c);
    if (s->damping != 1.0) for (long i = 0; i < N; i++) c[i] *= s->damping;
  } else {
    orc_blockjacobi_apply(s->lev[l], r, c, s->damping, 0);
  }
}
static void apply_smoother(vstate* s, int l, int steps, double* x, double* r) {
  long N = orc_mesh_dimension(s->lev[l]);
  memset(s->tmp1, 0, sizeof(double) * N);                  /* multigrid_impl.hh:73-74 */
  memset(s->tmp2, 0, sizeof(double) * N);
  for (int i = 0; i < steps; i++) {                        /* :76-81 */
    /* NOTE: with block-GS, tmp1 is NOT re-zeroed between steps in the reference (iterate()
     * continues from the previous tmp1); reproduced as is. */
    lvl_smooth(s, l, s->tmp1, r);
    for (long k = 0; k < N; k++) x[k] += s->tmp1[k];
    lvl_apply(s, l, s->tmp1, s->tmp2);
    for (long k = 0; k < N; k++) r[k] -= s->tmp2[k];
  }
}
static void apply_level(vstate* s, int l) {
  double* x = s->x[l]; double* r = s->r[l];
  long N = orc_mesh_dimension(s->lev[l]);
  if (l == 0) {                                            /* :93-96, coarse solver solversetup.hh:198-215 */
    if (s->smoother == 0) {
      for (int i = 0; i < s->coarse_its; i++) orc_blockgs_iterate(s->mats[0], r, x);
    } else {
      /* coarse_its damped block-Jacobi iterations x += w D^-1 (r - A x) from x = 0 */
      double* t = (double*)malloc(sizeof(double) * N * 2); double* res = t + N;
      for (int i = 0; i < s->coarse_its; i++) {
        lvl_apply(s, 0, x, t);
        for (long k = 0; k < N; k++) res[k] = r[k] - t[k];
        orc_blockjacobi_apply(s->lev[0], res, t, s->damping, 0);
        for (long k = 0; k < N; k++) x[k] += t[k];
      }
      free(t);
    }
    return;
  }
  apply_smoother(s, l, s->pre, x, r);                      /* :99 */
  orc_restrict(s->lev[l], s->lev[l - 1], r, s->r[l - 1]);  /* :103 */
  memset(s->x[l - 1], 0, sizeof(double) * orc_mesh_dimension(s->lev[l - 1]));
  apply_level(s, l - 1);                                   /* mu_ = 1 (multigrid.hh:67) */
  orc_prolong(s->lev[l], s->lev[l - 1], s->x[l - 1], s->tmp1);   /* :108 */
  for (long k = 0; k < N; k++) x[k] += s->tmp1[k];
  memset(s->tmp2, 0, sizeof(double) * N);
  lvl_apply(s, l, s->tmp1, s->tmp2);
  for (long k = 0; k < N; k++) r[k] -= s->tmp2[k];
  apply_smoother(s, l, s->post, x, r);                     /* :116 */
}

void orc_vcycle(int nlev, omesh* const* levels, obcrs* const* mats, int smoother, double damping,
                int pre, int post, int coarse_its, double* x, double* b) {
  vstate s = {nlev, levels, mats, smoother, damping, pre, post, coarse_its, NULL, NULL, NULL, NULL};
  int fine = nlev - 1;
  long N = orc_mesh_dimension(levels[fine]);
  s.x = (double**)malloc(sizeof(double*) * nlev * 2); s.r = s.x + nlev;
  for (int l = 0; l < nlev; l++) {
    long n = orc_mesh_dimension(levels[l]);
    s.x[l] = (double*)calloc(n, sizeof(double)); s.r[l] = (double*)calloc(n, sizeof(double));
  }
  s.tmp1 = (double*)calloc(N, sizeof(double)); s.tmp2 = (double*)calloc(N, sizeof(double));
  lvl_apply(&s, fine, x, s.tmp1);                          /* multigrid_impl.hh:30-36 */
  for (long k = 0; k < N; k++) s.r[fine][k] = b[k] - s.tmp1[k];
  apply_level(&s, fine);
  for (long k = 0; k < N; k++) { x[k] += s.x[fine][k]; b[k] = s.r[fine][k]; }   /* :60-61 */
  for (int l = 0; l < nlev; l++) { free(s.x[l]); free(s.r[l]); }
  free(s.x); free(s.tmp1); free(s.tmp2);
}

/* ------------------------------------------------------------------------ */
/* helpers                                                                   */
/* ------------------------------------------------------------------------ */
/* test/randomvector.hh:11-21: std::mt19937 seeded with `seed`, std::normal_distribution<>(0,1)
 * in block order.  Restated for libstdc++ (the reference's CI toolchains): mt19937 per the
 * standard; generate_canonical<double,53> draws two 32-bit words; normal_distribution is the
 * Marsaglia polar method returning y*mult first and caching x*mult. */
typedef struct { uint32_t mt[624]; int idx; } mt_t;
static void mt_seed(mt_t* g, uint32_t s) {
  g->mt[0] = s;
  for (int i = 1; i < 624; i++) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
static uint32_t mt_next(mt_t* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; i++) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
static double mt_canonical(mt_t* g) {
  /* libstdc++ generate_canonical<double,53,mt19937>: k = 2 draws, range 2^32 */
  double lo = (double)mt_next(g);
  double hi = (double)mt_next(g);
  double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}
void orc_fill_random(double* v, long n, unsigned seed) {
  mt_t g; mt_seed(&g, seed);
  int have = 0; double saved = 0;
  for (long i = 0; i < n; i++) {
    if (have) { have = 0; v[i] = saved; continue; }
    double x, y, r2;
    do {
      x = 2.0 * mt_canonical(&g) - 1.0;
      y = 2.0 * mt_canonical(&g) - 1.0;
      r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0.0);
    double mult = sqrt(-2.0 * log(r2) / r2);
    saved = x * mult; have = 1;
    v[i] = y * mult;
  }
}
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
