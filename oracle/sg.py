"""Oracle: the stochastic-Galerkin data model, operator application and PCG.

Restates
  * ``MultiVector``                 reference src/spuq/application/egsz/multi_vector.py:29-166
  * ``ListCoefficientField``        reference src/spuq/application/egsz/coefficient_field.py:55-104
  * ``MultiOperator.apply``         reference src/spuq/application/egsz/multi_operator2.py:38-84
  * ``PreconditioningOperator``     reference src/spuq/application/egsz/multi_operator2.py:139-165
  * ``pcg``                         reference src/spuq/application/egsz/pcg.py:15-53
  * ``ForgetfulVector``             reference src/spuq/utils/forgetful_vector.py:46-66
  * ``neighbour_combination``       reference src/spuq/application/egsz/residual_estimator2.py:107-131
  * ``refine_y``                    reference src/spuq/application/egsz/marking2.py:205-208
  * ``sample_realization``, ``compute_parametric_sample_solution``
                                    reference src/spuq/application/egsz/coefficient_field.py:45-52,
                                    src/spuq/application/egsz/sampling.py:59-75
  * ``prepare_rhs`` / ``pcg_solve`` reference src/spuq/application/egsz/adaptive_solver2.py:39-85,
                                    ``error_norm`` src/spuq/fem/fenics/fenics_utils.py:140-159
  * batched form / Delta matrices   reference src/spuq/linalg/tensor_operator.py:65-79,
                                    src/spuq/polyquad/structure_coefficients.py:11-41
Test infrastructure only (see oracle/__init__.py).
"""
import logging

import numpy as np
import scipy.sparse as sps

from .linalg import Operator, Scalar, Vector, inner
from .multiindex import Multiindex

logger = logging.getLogger(__name__)


class MultiVector(Vector):
    """dict {Multiindex -> Vector}; multi_vector.py:29-166."""

    def __init__(self, on_modify=lambda: None, multivector=None):
        self.mi2vec = dict()
        self.on_modify = on_modify
        if multivector is not None:
            for mu, vec in multivector.items():
                self[mu] = vec

    @property
    def max_order(self):
        return max(len(mu) for mu in self.keys())

    def __getitem__(self, mi):
        return self.mi2vec[mi]

    def __setitem__(self, mi, val):
        self.on_modify()
        self.mi2vec[mi] = val

    def __len__(self):
        return len(self.mi2vec)

    def keys(self):
        return self.mi2vec.keys()

    def values(self):
        return self.mi2vec.values()

    def items(self):
        return self.mi2vec.items()

    iteritems = items

    def active_indices(self):
        return sorted(self.keys())

    def copy(self):
        mv = type(self)()
        for mi in self.keys():
            mv[mi] = self[mi].copy()
        return mv

    def set_defaults(self, multiindex_set, init_vector):
        self.on_modify()
        for mi in multiindex_set:
            self[Multiindex(mi)] = init_vector.copy()

    def set_zero(self):
        for mu in self.active_indices():
            self[mu].set_zero()

    def __eq__(self, other):
        return type(self) == type(other) and self.mi2vec == other.mi2vec

    __hash__ = None

    def __neg__(self):
        new = self.copy()
        for mi in self.active_indices():
            new[mi] = -self[mi]
        return new

    def __iadd__(self, other):
        assert self.active_indices() == other.active_indices()
        self.on_modify()
        for mi in self.active_indices():
            self[mi] += other[mi]
        return self

    def __isub__(self, other):
        assert self.active_indices() == other.active_indices()
        self.on_modify()
        for mi in self.active_indices():
            self[mi] -= other[mi]
        return self

    def __imul__(self, other):
        assert isinstance(other, Scalar)
        self.on_modify()
        for mi in self.keys():
            self[mi] *= other
        return self

    def __inner__(self, other):
        assert isinstance(other, MultiVector)
        s = 0.0
        for mi in self.keys():
            s += inner(self[mi], other[mi])
        return s

    def __repr__(self):
        return "<%s keys=%s>" % (type(self).__name__, list(self.mi2vec.keys()))

    # slab view used by the tests: columns in active_indices() order (multi_vector.py:388-403)
    def to_slab(self):
        return np.stack([self[mu].coeffs for mu in self.active_indices()], axis=1)


class ListCoefficientField:
    """(mean_func, [(a_m, rv_m)]); coefficient_field.py:55-104."""

    def __init__(self, mean_func, funcs, rvs):
        assert len(funcs) == len(rvs)
        self._mean_func, self._funcs, self._rvs = mean_func, list(funcs), list(rvs)

    @classmethod
    def create_with_iid_rvs(cls, mean_func, funcs, rv):
        return cls(mean_func, funcs, [rv] * len(funcs))

    mean_func = property(lambda self: self._mean_func)
    funcs = property(lambda self: self._funcs)
    rvs = property(lambda self: self._rvs)

    def __len__(self):
        return len(self._funcs)

    def __getitem__(self, i):
        assert i < len(self._funcs), "invalid index"
        return self._funcs[i], self._rvs[i]

    def sample_realization(self, Lambda, RV_samples):
        """sample_map[mu] = prod_m P_{mu_m}(y_m) over the stored part of mu; coefficient_field.py:45-52."""
        sample_map = {}
        for mu in Lambda:
            univariate_vals = [self[m][1].orth_polys[mu[m]](RV_samples[m]) for m in range(len(mu))]
            sample_map[mu] = float(np.prod(univariate_vals)) if univariate_vals else 1.0
        return sample_map, RV_samples


class MultiOperator(Operator):
    """The SG operator on one shared spatial basis; multi_operator2.py:26-84."""

    def __init__(self, coeff_field, assemble_0, assemble_m=None, domain=None, codomain=None):
        if not (isinstance(coeff_field, ListCoefficientField) or hasattr(coeff_field, "mean_func")):
            raise TypeError("coeff_field")
        if not callable(assemble_0) or not (assemble_m is None or callable(assemble_m)):
            raise TypeError("assemble callback")
        self._assemble_0 = assemble_0
        self._assemble_m = assemble_m or assemble_0
        self._coeff_field = coeff_field
        self._domain, self._codomain = domain, codomain

    domain = property(lambda self: self._domain)
    codomain = property(lambda self: self._codomain)

    def apply(self, w, only=None):
        """``only`` (not in the reference): restrict the outer loop to these output indices; the
        loop body is unchanged.  Used to time / check a sample of columns of a large problem."""
        if not isinstance(w, MultiVector):
            raise TypeError("MultiOperator.apply needs a MultiVector")
        v = 0 * w if only is None else MultiVector()                  # :42
        Lambda = w.active_indices()                                   # :43
        maxm = w.max_order                                            # :44
        if len(self._coeff_field) < maxm:                             # :45-48
            logger.warning("insufficient length of coefficient field for MultiVector "
                           "(%i instead of %i", len(self._coeff_field), maxm)
            maxm = len(self._coeff_field)
        basis = w[Lambda[0]].basis                                    # :51 (shared basis)
        for mu in (Lambda if only is None else only):                 # :52
            A0 = self._assemble_0(basis, self._coeff_field.mean_func)
            cur_v = A0 * w[mu]                                        # :56
            for m in range(maxm):                                     # :59
                am_f, am_rv = self._coeff_field[m]
                Am = self._assemble_m(basis, am_f)
                beta = am_rv.orth_polys.get_beta(mu[m])               # :66
                cur_w = -beta[0] * w[mu]                              # :69
                mu1 = mu.inc(m)
                if mu1 in Lambda:                                     # :72-74
                    cur_w += beta[1] * w[mu1]
                mu2 = mu.dec(m)
                if mu2 in Lambda:                                     # :77-79
                    cur_w += beta[-1] * w[mu2]
                cur_v += Am * cur_w                                   # :82
            v[mu] = cur_v
        return v


class PreconditioningOperator(Operator):
    """Mean-based block-diagonal preconditioner; multi_operator2.py:139-165."""

    def __init__(self, mean_func, assemble_solver, domain=None, codomain=None):
        self._assemble_solver = assemble_solver
        self._mean_func = mean_func
        self._domain, self._codomain = domain, codomain

    domain = property(lambda self: self._domain)
    codomain = property(lambda self: self._codomain)

    def apply(self, w):
        v = 0 * w
        for mu in w.active_indices():
            A0 = self._assemble_solver(w[mu].basis, self._mean_func)
            v[mu] = A0 * w[mu]
        return v


class ForgetfulVector:
    """Ring buffer remembering the last ``remember`` entries; forgetful_vector.py:46-66."""

    def __init__(self, remember, start=0):
        self.rem, self.curr = remember, start
        self.items = remember * [None]

    def __getitem__(self, k):
        if not self.curr - self.rem < k <= self.curr:
            raise IndexError("index out of range: %d not in [%d, %d]"
                             % (k, self.curr - self.rem + 1, self.curr))
        return self.items[self.curr - k]

    def __setitem__(self, k, item):
        if not self.curr - self.rem < k <= self.curr + 1:
            raise IndexError("index out of range: %d not in [%d, %d]"
                             % (k, self.curr - self.rem + 1, self.curr + 1))
        if k == self.curr + 1:
            self.items = [item] + self.items[:-1]
            self.curr += 1
        else:
            self.items[self.curr - k] = item


def pcg(A, f, P, w0, eps=1e-4, maxiter=100, history=None):
    """Preconditioned CG; pcg.py:15-53.  Returns (w, zeta, i).

    ``history`` (a list, optional, not in the reference) receives zeta of every pass so tests can
    compare trajectories at a fixed iteration index.
    """
    w, rho, s, v, z, alpha = (ForgetfulVector(1) for _ in range(6))
    zeta = ForgetfulVector(2)
    w[0] = w0
    rho[0] = f - A * w[0]
    s[0] = P * rho[0]
    v[0] = s[0]
    zeta[0] = inner(rho[0], s[0])
    for i in range(1, maxiter):
        if history is not None:
            history.append(float(zeta[i - 1]))
        if logger.isEnabledFor(logging.INFO):
            logger.info("pcg iter: %s -> zeta=%s, rho^2=%s", i, zeta[i - 1],
                        inner(rho[i - 1], rho[i - 1]))
        if zeta[i - 1] < 0:
            raise Exception("Preconditioner for PCG is not positive definite (%s)" % zeta[i - 1])
        if zeta[i - 1] <= eps ** 2:
            return (w[i - 1], zeta[i - 1], i)
        z[i - 1] = A * v[i - 1]
        alpha[i - 1] = inner(z[i - 1], v[i - 1])
        if alpha[i - 1] == 0:
            raise Exception("Matrix for PCG is singular (%s)" % alpha[i - 1])
        elif alpha[i - 1] < 0:
            raise Exception("Matrix for PCG is not positive definite (%s)" % alpha[i - 1])
        w[i] = w[i - 1] + zeta[i - 1] / alpha[i - 1] * v[i - 1]
        rho[i] = rho[i - 1] - zeta[i - 1] / alpha[i - 1] * z[i - 1]
        s[i] = P * rho[i]
        zeta[i] = inner(rho[i], s[i])
        v[i] = s[i] + zeta[i] / zeta[i - 1] * v[i - 1]
    raise Exception("PCG did not converge")


# ---- the direct caller of pcg: right-hand side and solve ------------------
def prepare_rhs(A, w, coeff_field, pde):
    """Right-hand side of the SG system; adaptive_solver2.py:39-68.

    ``pde`` is the discretisation object of the reference (``FEMDiscretisation``,
    fem_discretisation.py:256-352) reduced to what this path calls: ``assemble_rhs(basis, coeff,
    withDirichletBC, withNeumannBC, f)`` returning the coefficient array and
    ``set_dirichlet_bc_entries(vec, homogeneous)``.  The zero load the reference builds as a FEniCS
    function (:45-50) is the scalar 0.0 here."""
    b = 0 * w                                                               # :40
    assert b.active_indices() == w.active_indices()                        # :41
    zero = Multiindex()
    b[zero].coeffs = pde.assemble_rhs(basis=b[zero].basis, coeff=coeff_field.mean_func,
                                      withNeumannBC=True)                  # :43-44
    zero_func = 0.0
    for m in range(w.max_order):                                           # :52
        eps_m = zero.inc(m)
        am_f, am_rv = coeff_field[m]
        beta = am_rv.orth_polys.get_beta(0)
        if eps_m in b.active_indices():                                    # :57
            g0 = b[eps_m].copy()
            g0.coeffs = pde.assemble_rhs(basis=b[eps_m].basis, coeff=am_f, withNeumannBC=False, f=zero_func)
            pde.set_dirichlet_bc_entries(g0, homogeneous=True)
            b[eps_m] += beta[1] * g0                                       # :61
        g0 = b[zero].copy()
        g0.coeffs = pde.assemble_rhs(basis=b[zero].basis, coeff=am_f, f=zero_func)
        pde.set_dirichlet_bc_entries(g0, homogeneous=True)
        b[zero] += beta[0] * g0                                            # :66
    return b


def error_norm(vec1, vec2, normstr, pde):
    """sqrt(sum_mu norm(vec1[mu] - vec2[mu])^2); fenics_utils.py:140-152.  ``normstr`` is a norm
    name handed to ``pde.norm`` (dolfin's ``norm`` in the reference) or a callable."""
    assert vec1.active_indices() == vec2.active_indices()
    e = 0.0
    for mi in vec1.keys():
        err = vec1[mi].coeffs - vec2[mi].coeffs
        e += (pde.norm(err, normstr) if isinstance(normstr, str) else normstr(err)) ** 2
    return float(np.sqrt(e))


def pcg_solve(A, w, coeff_field, pde, stats, pcg_eps, pcg_maxiter):
    """adaptive_solver2.py:71-85: right-hand side, mean-based preconditioner, PCG from w, residual
    of the returned iterate in two norms."""
    b = prepare_rhs(A, w, coeff_field, pde)
    P = PreconditioningOperator(coeff_field.mean_func, pde.assemble_solve_operator)
    w, zeta, numit = pcg(A, b, P, w0=w, eps=pcg_eps, maxiter=pcg_maxiter)
    b2 = A * w                                                             # :79
    stats["RESIDUAL-L2"] = error_norm(b, b2, "L2", pde)
    stats["RESIDUAL-H1A"] = error_norm(b, b2, pde.energy_norm, pde)
    stats["DOFS"] = sum(b[mu].basis.dim for mu in b.keys())
    stats["CELLS"] = sum(pde.num_cells(b[mu].basis) for mu in b.keys())
    return w, zeta


def neighbour_combination(w, coeff_field, m):
    """res[mu] = -beta[0] w[mu] + beta[1] w[mu+e_m] + beta[-1] w[mu-e_m] for every active mu;
    residual_estimator2.py:107-131 (the loop body for one m, before the FEM forms are applied)."""
    Lambda = w.active_indices()
    out = MultiVector()
    am_f, am_rv = coeff_field[m]
    for mu in Lambda:
        beta = am_rv.orth_polys.get_beta(mu[m])
        res = -beta[0] * w[mu]
        mu1 = mu.inc(m)
        if mu1 in Lambda:
            res += beta[1] * w[mu1]
        mu2 = mu.dec(m)
        if mu2 in Lambda:
            res += beta[-1] * w[mu2]
        out[mu] = res
    return out


def supp(Lambda):
    """Union of the supports of the indices of Lambda; multi_vector.py:24-26."""
    return set.union(*[set(int(i) for i in np.flatnonzero(mu.as_array)) for mu in Lambda])


def evaluate_upper_tail_bound(w, coeff_field, energynorm, ainfty, add_maxm=10):
    """Upper tail bounds of EGSZ section 3.2; residual_estimator2.py:176-290 with the two FEM ingredients
    supplied by the caller: ``energynorm(vec)`` (pde.energy_norm applied to w[mu], :207-211) and
    ``ainfty(m)`` = ||a_m / a_mean||_inf (:181-205).  M = w.max_order + add_maxm (:271), capped at the
    length of the coefficient field when that is finite (the reference's own commented-out variant, :270).
    Returns (global_zeta, zeta, zeta_bar, eval_zeta_m) like the reference; note that LambdaBoundary
    (:213-223) yields a boundary index once per neighbour it has in Lambda and :279 ADDS eval_zeta for
    every yield."""
    from collections import defaultdict
    from math import sqrt

    def LambdaBoundary(Lambda):
        suppLambda = supp(Lambda)
        for mu in Lambda:
            for m in suppLambda:
                mu1 = mu.inc(m)
                if mu1 not in Lambda:
                    yield mu1
                mu2 = mu.dec(m)
                if mu2 not in Lambda and mu2 is not None:
                    yield mu2

    def eval_zeta_bar(mu, suppLambda, normw, M):
        zz = 0
        for m in range(M):
            if m in suppLambda:
                continue
            _, am_rv = coeff_field[m]
            beta = am_rv.orth_polys.get_beta(mu[m])
            zz += (beta[1] * ainfty(m)) ** 2
        return normw[mu] * sqrt(zz)

    def eval_zeta(mu, Lambda, normw, M=None, this_m=None):
        z = 0
        if this_m is None:
            for m in range(M):
                _, am_rv = coeff_field[m]
                beta = am_rv.orth_polys.get_beta(mu[m])
                a = ainfty(m)
                mu1 = mu.inc(m)
                if mu1 in Lambda:
                    z += a * beta[1] * normw[mu1]
                mu2 = mu.dec(m)
                if mu2 in Lambda:
                    z += a * beta[-1] * normw[mu2]
            return z
        m = this_m
        _, am_rv = coeff_field[m]
        beta = am_rv.orth_polys.get_beta(mu[m])
        return ainfty(m) * beta[1] * normw[mu]

    Lambda = w.active_indices()
    suppLambda = supp(Lambda)
    M = w.max_order + add_maxm
    try:
        M = min(M, len(coeff_field))
    except TypeError:
        pass
    normw = {mu: energynorm(w[mu]) for mu in Lambda}
    zeta = defaultdict(int)
    for nu in LambdaBoundary(Lambda):
        zeta[nu] += eval_zeta(nu, Lambda, normw, M)
    zeta_bar = {mu: eval_zeta_bar(mu, suppLambda, normw, M) for mu in Lambda}
    global_zeta = sqrt(sum(v ** 2 for v in zeta.values()) + sum(v ** 2 for v in zeta_bar.values()))
    return global_zeta, zeta, zeta_bar, (lambda mu, m: eval_zeta(mu, Lambda, normw, M, m))


def refine_multivector(w, cell_ids=None):
    """MultiVectorSharedBasis.refine, multi_vector.py:474-479: every vector prolongated onto the refined
    (shared) basis."""
    _, prolongate, _ = w[w.active_indices()[0]].basis.refine(cell_ids)
    mv = MultiVector()
    for mi in w.keys():
        mv[mi] = prolongate(w[mi])
    return mv


def refine_y(w, new_mi):
    """Marking.refine_y, marking2.py:205-208: a zero vector for every new multi-index."""
    some = w[w.active_indices()[0]]
    for mu in new_mi:
        z = some.copy()
        z.coeffs = np.zeros_like(some.coeffs)
        w[mu] = z


def compute_parametric_sample_solution(RV_samples, coeff_field, w, proj_basis=None, cache=None):
    """The expansion evaluated at one parameter sample, sum_mu w[mu] * prod_m P_{mu_m}(y_m);
    sampling.py:59-75 with all vectors already on the projection basis (shared basis: the projection
    of :46-52 is the identity)."""
    Lambda = w.active_indices()
    sample_map, _ = coeff_field.sample_realization(Lambda, RV_samples)
    return sum(w[mu] * sample_map[mu] for mu in Lambda)


# ---- structural helpers used by the parity tests -----------------------------------------------
def neighbour_tables(Lambda, maxm):
    """Positions of mu+e_m / mu-e_m inside the sorted list ``Lambda`` (-1 when absent), found the
    way the reference finds them: ``mu.inc(m) in Lambda`` / ``mu.dec(m) in Lambda``
    (multi_operator2.py:72-79), i.e. by Multiindex equality, not by any array trick."""
    pos = {mu: k for k, mu in enumerate(Lambda)}
    up = -np.ones((maxm, len(Lambda)), dtype=np.int32)
    dn = -np.ones((maxm, len(Lambda)), dtype=np.int32)
    for k, mu in enumerate(Lambda):
        for m in range(maxm):
            mu1 = mu.inc(m)
            if mu1 in pos:
                up[m, k] = pos[mu1]
            mu2 = mu.dec(m)
            if mu2 is not None and mu2 in pos:
                dn[m, k] = pos[mu2]
    return up, dn


def stochastic_matrices(Lambda, rvs, maxm):
    """K_m (L x L, CSR) with the row-degree convention of MultiOperator.apply:
    K_m[mu, mu] = -beta0(mu_m), K_m[mu, mu+e_m] = beta1(mu_m), K_m[mu, mu-e_m] = beta_{-1}(mu_m)."""
    L = len(Lambda)
    up, dn = neighbour_tables(Lambda, maxm)
    mats = []
    for m in range(maxm):
        polys = rvs[m].orth_polys
        K = sps.lil_matrix((L, L))
        for k, mu in enumerate(Lambda):
            b = polys.get_beta(mu[m])
            if b[0] != 0.0:
                K[k, k] = -b[0]
            if up[m, k] >= 0:
                K[k, up[m, k]] = b[1]
            if dn[m, k] >= 0:
                K[k, dn[m, k]] = b[-1]
        mats.append(K.tocsr())
    return mats


def apply_batched(A0, Ams, Ks, W):
    """Batched CPU form  Y = A0 W + sum_m A_m (W K_m^T)  on the N x L slab W (columns in sorted
    Lambda order).  This is the reference's TensorOperator formulation (tensor_operator.py:65-79,
    tensor_vector.py:50-56) with the MultiOperator sign / degree conventions; it is the form the GPU
    implements and the 'best effort' CPU baseline of BASELINE.md."""
    Y = A0 @ W
    for Am, K in zip(Ams, Ks):
        if K.nnz:
            Y += Am @ (K @ W.T).T
    return Y


def evaluate_triples(polysys, I_a, I_b):
    """structure_coefficients.py:11-41: L[0] = identity pattern, L[k+1][i,j] = beta_k(mu_j[k])[d]
    with d = mu_i[k]-mu_j[k] in {-1,0,1} when mu_i and mu_j agree off component k."""
    Mi, Nj, K = I_a.shape[0], I_b.shape[0], I_a.shape[1]
    if not isinstance(polysys, (list, tuple)):
        polysys = [polysys] * K
    out = [sps.dok_matrix((Mi, Nj)) for _ in range(-1, K)]
    for k in range(-1, K):
        for i, mui in enumerate(I_a):
            for j, muj in enumerate(I_b):
                same = mui == muj
                if np.all(same[:k]) and np.all(same[k + 1:]):
                    delta = int(mui[k]) - int(muj[k])
                    if -1 <= delta <= 1:
                        val = polysys[k].get_beta(int(muj[k]))[delta] if k != -1 else 1
                        if val != 0.0:
                            out[k + 1][i, j] = val
    return [m.asformat("csr") for m in out]
