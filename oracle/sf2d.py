"""TEST INFRASTRUCTURE (oracle) -- a third, independent CPU formulation of the 2-D SIPG operator apply: the reference's
sum-factorised local operator with Gauss-LOBATTO quadrature,

    /root/reference/dune/hpdg/matrix-free/localoperators/sfipdg.hh
        bind            :61-80     rule of order 2p - 1 + 1 = GL with p + 2 points (getRule :395-406)
        computeBulk     :111-166   B^T U L tensor evaluation of the reference gradient, Jacobian scaling, C += A X^T B^T
        computeFace     :168-326   once per face from the element with the larger index (:196), u_diff / du_sum / X (:233-277),
                                   -{d_n phi}[u] (computeDPhi :553-624), -{d_n u}[phi] (:281-297), penalty (:300-313),
                                   outer rows scattered (:315-324)
        Dirichlet edge  :329-393   weight 1 instead of 1/2, penalty sigma p^2 / |e|
        outerBind       :421-469   hp faces: both sides are evaluated at the rule of the HIGHER degree
    .../gausslobattomatrices.hh:29-57,77-107   l_i(xi_q), l_i'(xi_q) by explicit Lagrange products on the sorted GL nodes

It shares NO code with oracle/hpdg_oracle.c (Gauss-Legendre quadrature loops, Newton node generator) nor with the product's
tables: nodes and weights come from numpy's Legendre module (companion-matrix roots of P'_{m-1}), the tensor contractions are
numpy einsums.  Only tests/ import it.  Structured axis-parallel mesh, elements numbered x-fastest, local index i0 + i1 (p+1),
faces 0/1 = x-/x+, 2/3 = y-/y+ (sfipdg.hh:496-516).
"""
import numpy as np
from numpy.polynomial import legendre as npleg


def gauss_lobatto(m):
    """m-point Gauss-Lobatto rule on [0, 1], ascending (the rule dune-geometry returns for order 2m - 3, sorted as in
    gausslobattomatrices.hh:33-35).  Interior nodes: roots of P'_{m-1}; weights 2 / (m (m-1) P_{m-1}(x)^2) on [-1, 1]."""
    if m == 1:
        return np.array([0.5]), np.array([1.0])
    pm1 = npleg.Legendre.basis(m - 1)
    xi = np.concatenate(([-1.0], np.sort(pm1.deriv().roots().real), [1.0])) if m > 2 else np.array([-1.0, 1.0])
    w = 2.0 / (m * (m - 1) * pm1(xi) ** 2)
    return 0.5 * (xi + 1.0), 0.5 * w


def lagrange_tables(p, pts):
    """values[i][q] = l_i(pts[q]), derivatives[i][q] = l_i'(pts[q]) for the degree-p Lagrange basis on the p+1 GL nodes
    (gausslobattomatrices.hh:46-57, 94-107: explicit products)."""
    nodes, _ = gauss_lobatto(p + 1) if p > 0 else (np.array([0.5]), None)
    n = p + 1
    V = np.ones((n, len(pts)))
    D = np.zeros((n, len(pts)))
    for i in range(n):
        for j in range(n):
            if j != i:
                V[i] *= (pts - nodes[j]) / (nodes[i] - nodes[j])
        for j in range(n):
            if j == i:
                continue
            prod = np.full(len(pts), 1.0 / (nodes[i] - nodes[j]))
            for l in range(n):
                if l != i and l != j:
                    prod *= (pts - nodes[l]) / (nodes[i] - nodes[l])
            D[i] += prod
    return V, D


class SumFactIPDG2D:
    def __init__(self, n, L=(1.0, 1.0), degree=1, sigma=2.0, dirichlet=True):
        self.n = tuple(int(v) for v in n)
        self.h = (L[0] / self.n[0], L[1] / self.n[1])
        ne = self.n[0] * self.n[1]
        self.deg = np.full(ne, degree, dtype=int) if np.isscalar(degree) else np.asarray(degree, dtype=int).copy()
        self.sigma, self.dirichlet = sigma, dirichlet
        self.off = np.concatenate(([0], np.cumsum((self.deg + 1) ** 2)))
        self.ndof = int(self.off[-1])
        self._cache = {}

    def _tab(self, p, m):   # (basis degree, rule size) -> (points, weights, values, derivatives)   (getMatrix, :408-419)
        key = (p, m)
        if key not in self._cache:
            x, w = gauss_lobatto(m)
            self._cache[key] = (x, w) + lagrange_tables(p, x)
        return self._cache[key]

    @staticmethod
    def _edge_coeffs(C, f):   # coefficientsOnEdge (:493-521); C[i1][i0]
        return C[:, 0] if f == 0 else C[:, -1] if f == 1 else C[0, :] if f == 2 else C[-1, :]

    def _edge_grad(self, C, p, m, f):
        """reference gradient (d/dxi0, d/dxi1) of the local function at the m rule points of edge f (computeDerivatives :640-703)"""
        _, _, V, D = self._tab(p, m)
        _, _, _, Dn = self._tab(p, p + 2)   # normal direction: the element's own table, end columns (:646,:659,:679,:692)
        end = 0 if f % 2 == 0 else -1
        if f < 2:   # edge x = const: tangential index i1
            d0 = (C @ Dn[:, end]) @ V          # sum_{i1} (sum_{i0} C[i1][i0] l_i0'(end)) l_i1(q)
            d1 = self._edge_coeffs(C, f) @ D
        else:       # edge y = const: tangential index i0
            d0 = self._edge_coeffs(C, f) @ D
            d1 = (Dn[:, end] @ C) @ V
        return d0, d1

    def _add_dphi(self, out, p, m, f, X, u):
        """out[i1][i0] += sum_q grad phi_{i0,i1}(q) . X[:, q] u[q]   (computeDPhi :553-624)"""
        _, _, V, D = self._tab(p, m)
        _, _, Vn, Dn = self._tab(p, p + 2)
        end = 0 if f % 2 == 0 else -1
        if f < 2:
            front = V @ (X[0] * u)     # tangential values, normal derivative of phi
            back = D @ (X[1] * u)      # tangential derivative, normal value
            out += np.outer(front, Dn[:, end]) + np.outer(back, Vn[:, end])
        else:
            front = V @ (X[1] * u)
            back = D @ (X[0] * u)
            out += np.outer(Dn[:, end], front) + np.outer(Vn[:, end], back)

    @staticmethod
    def _add_edge(out, vals, f):   # addOnEdge (:524-550)
        if f == 0:
            out[:, 0] += vals
        elif f == 1:
            out[:, -1] += vals
        elif f == 2:
            out[0, :] += vals
        else:
            out[-1, :] += vals

    def apply(self, x, factor=1.0):
        n0, n1 = self.n
        hx, hy = self.h
        y = np.zeros(self.ndof)
        blk = lambda v, e: v[self.off[e]:self.off[e + 1]].reshape(self.deg[e] + 1, self.deg[e] + 1)   # [i1][i0]
        for e in range(n0 * n1):
            p = int(self.deg[e])
            ex, ey = e % n0, e // n0
            C = blk(x, e)
            loc = np.zeros_like(C)
            # ---- computeBulk (:111-166) ----
            xq, wq, V, D = self._tab(p, p + 2)
            dx = np.einsum('ba,aq,br->rq', C, D, V)   # d/dxi0 u at (q0 = q, q1 = r), stored [q1][q0]
            dy = np.einsum('ba,aq,br->rq', C, V, D)
            W = np.outer(wq, wq) * hx * hy            # gamma * w_q0 * w_q1
            X0, X1 = dx * W / hx ** 2, dy * W / hy ** 2   # J^-1 J^-T grad (:131-137)
            loc += np.einsum('rq,aq,br->ba', X0, D, V) + np.einsum('rq,aq,br->ba', X1, V, D)
            # ---- computeFace (:168-326) ----
            for f in range(4):
                d, s = f // 2, f % 2
                nb = (ex + (1 if s else -1), ey) if d == 0 else (ex, ey + (1 if s else -1))
                normal = np.zeros(2)
                normal[d] = 1.0 if s else -1.0
                flen = hy if d == 0 else hx           # |e| = is.geometry().volume()
                jinv = np.array([1.0 / hx, 1.0 / hy])
                if not (0 <= nb[0] < n0 and 0 <= nb[1] < n1):
                    if not self.dirichlet:
                        continue
                    # computeDirichletBoundaryEdge (:329-393)
                    m = p + 2
                    xq, wq, V, D = self._tab(p, m)
                    pen = self.sigma * p ** 2 / flen
                    u = self._edge_coeffs(C, f) @ V
                    d0, d1 = self._edge_grad(C, p, m, f)
                    fac = wq * flen
                    u = u * fac
                    du = (d0 * jinv[0] * normal[0] + d1 * jinv[1] * normal[1]) * fac
                    Xn = -np.outer(jinv * normal, np.ones(m))
                    self._add_dphi(loc, p, m, f, Xn, u)
                    self._add_edge(loc, -(V @ du), f)
                    self._add_edge(loc, V @ (u * pen), f)
                    continue
                o = nb[0] + n0 * nb[1]
                if e < o:      # the element with the larger index does the work (:196)
                    continue
                po = int(self.deg[o])
                fo = f ^ 1
                m = max(p, po) + 2                       # outerBind (:421-469): the rule of the higher degree
                xq, wq, V, D = self._tab(p, m)
                _, _, Vo, Do = self._tab(po, m)
                Co = blk(x, o)
                oloc = np.zeros_like(Co)
                pen = self.sigma * max(p, po) ** 2 / flen
                u = self._edge_coeffs(C, f) @ V - self._edge_coeffs(Co, fo) @ Vo
                d0i, d1i = self._edge_grad(C, p, m, f)
                d0o, d1o = self._edge_grad(Co, po, m, fo)
                fac = wq * flen
                u = u * fac
                du = ((d0i + d0o) * jinv[0] * normal[0] + (d1i + d1o) * jinv[1] * normal[1]) * fac * 0.5
                Xn = -0.5 * np.outer(jinv * normal, np.ones(m))
                self._add_dphi(loc, p, m, f, Xn, u)
                self._add_dphi(oloc, po, m, fo, Xn, u)
                self._add_edge(loc, -(V @ du), f)
                self._add_edge(oloc, Vo @ du, fo)
                self._add_edge(loc, V @ (u * pen), f)
                self._add_edge(oloc, -(Vo @ (u * pen)), fo)
                y[self.off[o]:self.off[o + 1]] += factor * oloc.ravel()
            y[self.off[e]:self.off[e + 1]] += factor * loc.ravel()
        return y


class RefinedSumFactIPDG2D:
    """The same operator on a NON-CONFORMING mesh: the base grid with the flagged cells split once into 2 x 2 children, hanging
    nodes on faces between a refined and an unrefined cell.  Follows the non-conforming branch of the reference's computeFace
    (sfipdg.hh:213-222: `if (not is.conforming())` the 1-D values / derivatives of BOTH sides are re-evaluated at the rule points
    mapped through geometryInInside / geometryInOutside, nonConformingMatrices :472-491; penalty sigma max(p)^2 / |intersection|
    :227-229; each side's own inverse Jacobian :251-252,262-270).  Leaf elements: base cell by base cell (x fastest), the four
    children of a refined cell x fastest -- the numbering of hpdg_create_refined_2d.  Intersections are found geometrically."""

    def __init__(self, n, refine, degree, L=(1.0, 1.0), sigma=2.0, dirichlet=True):
        self.n = tuple(int(v) for v in n)
        hx, hy = L[0] / self.n[0], L[1] / self.n[1]
        refine = np.asarray(refine).astype(bool).ravel()
        self.leaves = []   # (x0, y0, hx, hy)
        for c in range(self.n[0] * self.n[1]):
            cx, cy = c % self.n[0], c // self.n[0]
            if not refine[c]:
                self.leaves.append((cx * hx, cy * hy, hx, hy))
            else:
                for b in range(2):
                    for a in range(2):
                        self.leaves.append((cx * hx + a * hx / 2, cy * hy + b * hy / 2, hx / 2, hy / 2))
        ne = len(self.leaves)
        self.deg = np.full(ne, degree, dtype=int) if np.isscalar(degree) else np.asarray(degree, dtype=int).copy()
        assert len(self.deg) == ne
        self.sigma, self.dirichlet, self.Ldom = sigma, dirichlet, (L[0], L[1])
        self.off = np.concatenate(([0], np.cumsum((self.deg + 1) ** 2)))
        self.ndof = int(self.off[-1])
        self._sf = SumFactIPDG2D((1, 1))   # table cache and the edge helpers
        # intersections (e, f, o, s0, s1): side f of e against the opposite side of o, overlap [s0, s1] in physical coordinates
        self.inter, self.bnd = [], []
        eps = 1e-12 * max(L)
        for e, (x0, y0, ex, ey) in enumerate(self.leaves):
            for f in range(4):
                d, s = f // 2, f % 2
                pos = (x0 + s * ex) if d == 0 else (y0 + s * ey)            # the side's normal coordinate
                t0, t1 = (y0, y0 + ey) if d == 0 else (x0, x0 + ex)         # its tangential extent
                if abs(pos - (self.Ldom[d] if s else 0.0)) < eps:
                    self.bnd.append((e, f))
                    continue
                for o, (u0, v0, ox, oy) in enumerate(self.leaves):
                    opos = (u0 + (1 - s) * ox) if d == 0 else (v0 + (1 - s) * oy)
                    if o == e or abs(opos - pos) > eps:
                        continue
                    q0, q1 = (v0, v0 + oy) if d == 0 else (u0, u0 + ox)
                    a, b = max(t0, q0), min(t1, q1)
                    if b - a > eps:
                        self.inter.append((e, f, o, a, b))

    def _tang(self, e, f, a, b, m):
        """1-D values / derivatives of element e's tangential basis at the m rule points of the intersection [a, b]
        (nonConformingMatrices, sfipdg.hh:472-491: rule points mapped into the element)"""
        x0, y0, ex, ey = self.leaves[e]
        t0, ht = (y0, ey) if f < 2 else (x0, ex)
        xq, wq = gauss_lobatto(m)
        tau = (a + (b - a) * xq - t0) / ht
        return lagrange_tables(int(self.deg[e]), tau) + (wq,)

    def _grad(self, C, p, f, V, D):
        _, _, _, Dn = self._sf._tab(p, p + 2)
        end = 0 if f % 2 == 0 else -1
        ec = SumFactIPDG2D._edge_coeffs(C, f)
        if f < 2:
            return (C @ Dn[:, end]) @ V, ec @ D
        return ec @ D, (Dn[:, end] @ C) @ V

    def _dphi(self, out, p, f, V, D, X, u):
        _, _, Vn, Dn = self._sf._tab(p, p + 2)
        end = 0 if f % 2 == 0 else -1
        if f < 2:
            out += np.outer(V @ (X[0] * u), Dn[:, end]) + np.outer(D @ (X[1] * u), Vn[:, end])
        else:
            out += np.outer(Dn[:, end], V @ (X[1] * u)) + np.outer(Vn[:, end], D @ (X[0] * u))

    def apply(self, x, factor=1.0):
        y = np.zeros(self.ndof)
        blk = lambda v, e: v[self.off[e]:self.off[e + 1]].reshape(self.deg[e] + 1, self.deg[e] + 1)
        for e, (x0, y0, hx, hy) in enumerate(self.leaves):   # computeBulk (:111-166)
            p = int(self.deg[e])
            C = blk(x, e)
            xq, wq, V, D = self._sf._tab(p, p + 2)
            dx = np.einsum('ba,aq,br->rq', C, D, V)
            dy = np.einsum('ba,aq,br->rq', C, V, D)
            W = np.outer(wq, wq) * hx * hy
            loc = np.einsum('rq,aq,br->ba', dx * W / hx ** 2, D, V) + np.einsum('rq,aq,br->ba', dy * W / hy ** 2, V, D)
            y[self.off[e]:self.off[e + 1]] += factor * loc.ravel()
        for e, f in self.bnd:                                  # computeDirichletBoundaryEdge (:329-393)
            if not self.dirichlet:
                continue
            x0, y0, hx, hy = self.leaves[e]
            p = int(self.deg[e])
            d, s = f // 2, f % 2
            flen = hy if d == 0 else hx
            nrm = np.zeros(2)
            nrm[d] = 1.0 if s else -1.0
            jinv = np.array([1.0 / hx, 1.0 / hy])
            xq, wq, V, D = self._sf._tab(p, p + 2)
            C = blk(x, e)
            loc = np.zeros_like(C)
            fac = wq * flen
            u = (SumFactIPDG2D._edge_coeffs(C, f) @ V) * fac
            d0, d1 = self._grad(C, p, f, V, D)
            du = (d0 * jinv[0] * nrm[0] + d1 * jinv[1] * nrm[1]) * fac
            self._dphi(loc, p, f, V, D, -np.outer(jinv * nrm, np.ones(len(wq))), u)
            SumFactIPDG2D._add_edge(loc, -(V @ du), f)
            SumFactIPDG2D._add_edge(loc, V @ (u * self.sigma * p ** 2 / flen), f)
            y[self.off[e]:self.off[e + 1]] += factor * loc.ravel()
        for e, f, o, a, b in self.inter:                       # computeFace (:168-326), once per intersection (:196)
            if e < o:
                continue
            p, po, fo = int(self.deg[e]), int(self.deg[o]), f ^ 1
            m = max(p, po) + 2
            V, D, wq = self._tang(e, f, a, b, m)
            Vo, Do, _ = self._tang(o, fo, a, b, m)
            d, s = f // 2, f % 2
            nrm = np.zeros(2)
            nrm[d] = 1.0 if s else -1.0
            ji = np.array([1.0 / self.leaves[e][2], 1.0 / self.leaves[e][3]])
            jo = np.array([1.0 / self.leaves[o][2], 1.0 / self.leaves[o][3]])
            flen = b - a
            C, Co = blk(x, e), blk(x, o)
            loc, oloc = np.zeros_like(C), np.zeros_like(Co)
            fac = wq * flen
            u = (SumFactIPDG2D._edge_coeffs(C, f) @ V - SumFactIPDG2D._edge_coeffs(Co, fo) @ Vo) * fac
            d0i, d1i = self._grad(C, p, f, V, D)
            d0o, d1o = self._grad(Co, po, fo, Vo, Do)
            # the tangential derivative tables are d/dtau: the chain rule of the sub-face map is already in tau's definition
            du = ((d0i * ji[0] + d0o * jo[0]) * nrm[0] + (d1i * ji[1] + d1o * jo[1]) * nrm[1]) * fac * 0.5
            ones = np.ones(len(wq))
            self._dphi(loc, p, f, V, D, -0.5 * np.outer(ji * nrm, ones), u)
            self._dphi(oloc, po, fo, Vo, Do, -0.5 * np.outer(jo * nrm, ones), u)
            pen = self.sigma * max(p, po) ** 2 / flen
            SumFactIPDG2D._add_edge(loc, -(V @ du) + V @ (u * pen), f)
            SumFactIPDG2D._add_edge(oloc, Vo @ du - Vo @ (u * pen), fo)
            y[self.off[e]:self.off[e + 1]] += factor * loc.ravel()
            y[self.off[o]:self.off[o + 1]] += factor * oloc.ravel()
        return y

    def interpolate(self, fn):
        """nodal interpolation of fn(x, y) (qkgllocalinterpolation.hh:56-74: point evaluation at the GL nodes)"""
        v = np.zeros(self.ndof)
        for e, (x0, y0, hx, hy) in enumerate(self.leaves):
            p = int(self.deg[e])
            nodes = gauss_lobatto(p + 1)[0] if p > 0 else np.array([0.5])
            X, Y = np.meshgrid(x0 + hx * nodes, y0 + hy * nodes)   # [i1][i0]
            v[self.off[e]:self.off[e + 1]] = fn(X, Y).ravel()
        return v
