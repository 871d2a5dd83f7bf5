"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so (the plain-C restatement of the reference's
partial-assembly hot path, oracle_pa.c).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

P_D = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
P_I = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build():
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.isfile(path):
            build()
        _LIB = C.CDLL(path)
    return _LIB


class OrcOperator(C.Structure):
    _fields_ = [("dim", C.c_int), ("d1d", C.c_int), ("q1d", C.c_int), ("is_mass", C.c_int),
                ("symmetric", C.c_int), ("ne", C.c_int64), ("ndofs", C.c_int),
                ("gather_map", C.c_void_p), ("offsets", C.c_void_p), ("indices", C.c_void_p),
                ("B", C.c_void_p), ("G", C.c_void_p), ("D", C.c_void_p),
                ("ess", C.c_void_p), ("ness", C.c_int),
                ("xe", C.c_void_p), ("ye", C.c_void_p), ("z", C.c_void_p)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def restriction_build(ne, dof, ndofs, elem_dof, dof_map):
    g = np.empty(ne * dof, np.int32)
    o = np.empty(ndofs + 1, np.int32)
    i = np.empty(ne * dof, np.int32)
    dm = _i32(dof_map) if dof_map is not None and len(dof_map) else None
    lib().orc_restriction_build(ne, dof, ndofs, _p(_i32(elem_dof)),
                                _p(dm) if dm is not None else None, _p(g), _p(o), _p(i))
    return g, o, i


def restriction_mult(ne, dof, gather_map, x):
    y = np.empty(ne * dof)
    lib().orc_restriction_mult(ne, dof, _p(_i32(gather_map)), _p(_f64(x)), _p(y))
    return y


def restriction_mult_transpose(ndofs, dof, offsets, indices, xe, y=None):
    add = 0 if y is None else 1
    out = np.empty(ndofs) if y is None else _f64(y).copy()
    lib().orc_restriction_mult_transpose(ndofs, dof, _p(_i32(offsets)), _p(_i32(indices)),
                                         _p(_f64(xe)), _p(out), add)
    return out


def restriction_abs_mult_transpose(ndofs, dof, offsets, indices, xe):
    out = np.empty(ndofs)
    lib().orc_restriction_abs_mult_transpose(ndofs, dof, _p(_i32(offsets)),
                                             _p(_i32(indices)), _p(_f64(xe)), _p(out))
    return out


def gauss_legendre(n):
    p, w = np.empty(n), np.empty(n)
    lib().orc_gauss_legendre(n, _p(p), _p(w))
    return p, w


def gauss_lobatto(n):
    p, w = np.empty(n), np.empty(n)
    lib().orc_gauss_lobatto(n, _p(p), _p(w))
    return p, w


def basis_tables(nodes, qpts):
    nd, nq = len(nodes), len(qpts)
    B, G = np.empty(nq * nd), np.empty(nq * nd)
    lib().orc_basis_tables(nd, _p(_f64(nodes)), nq, _p(_f64(qpts)), _p(B), _p(G))
    return B, G


def tensor_weights(dim, w1):
    nq = len(w1)
    W = np.empty(nq ** dim)
    lib().orc_tensor_weights(dim, nq, _p(_f64(w1)), _p(W))
    return W


def geometric_factors(dim, nd1d, q1d, ne, Bn, Gn, nodesE):
    nq = q1d ** dim
    J = np.empty(nq * dim * dim * ne)
    detJ = np.empty(nq * ne)
    lib().orc_geometric_factors(dim, nd1d, q1d, C.c_int64(ne), _p(_f64(Bn)), _p(_f64(Gn)),
                                _p(_f64(nodesE)), _p(J), _p(detJ))
    return J, detJ


def diffusion_setup(dim, q1d, coeff_dim, ne, W, J, coeff):
    coeff = _f64(coeff)
    const = int(coeff.size == coeff_dim)
    sym = coeff_dim != dim * dim
    ncomp = (dim * (dim + 1) // 2) if sym else dim * dim
    D = np.empty(q1d ** dim * ncomp * ne)
    lib().orc_diffusion_setup(dim, q1d, coeff_dim, const, C.c_int64(ne), _p(_f64(W)),
                              _p(_f64(J)), _p(coeff), _p(D))
    return D


def mass_setup(nq, ne, W, detJ, coeff):
    coeff = _f64(coeff)
    D = np.empty(nq * ne)
    lib().orc_mass_setup(nq, C.c_int64(ne), _p(_f64(W)), _p(_f64(detJ)), _p(coeff),
                         int(coeff.size == 1), _p(D))
    return D


def diffusion_apply(dim, d1d, q1d, ne, B, G, D, X, symmetric=1):
    Y = np.zeros(d1d ** dim * ne)
    lib().orc_diffusion_apply(dim, d1d, q1d, C.c_int64(ne), symmetric, _p(_f64(B)),
                              _p(_f64(G)), _p(_f64(D)), _p(_f64(X)), _p(Y))
    return Y


def mass_apply(dim, d1d, q1d, ne, B, D, X):
    Y = np.zeros(d1d ** dim * ne)
    lib().orc_mass_apply(dim, d1d, q1d, C.c_int64(ne), _p(_f64(B)), _p(_f64(D)),
                         _p(_f64(X)), _p(Y))
    return Y


def diffusion_diag(dim, d1d, q1d, ne, B, G, D, symmetric=1):
    Y = np.zeros(d1d ** dim * ne)
    lib().orc_diffusion_diag(dim, d1d, q1d, C.c_int64(ne), symmetric, _p(_f64(B)),
                             _p(_f64(G)), _p(_f64(D)), _p(Y))
    return Y


def mass_diag(dim, d1d, q1d, ne, B, D):
    Y = np.zeros(d1d ** dim * ne)
    lib().orc_mass_diag(dim, d1d, q1d, C.c_int64(ne), _p(_f64(B)), _p(_f64(D)), _p(Y))
    return Y


def domain_lf(dim, d1d, q1d, vdim, ne, B, W, detJ, coeff, markers=None):
    """DomainLFIntegrator::AssembleDevice on a zeroed E-vector (D1D^dim, vdim, NE)."""
    coeff = _f64(coeff)
    y = np.zeros(ne * vdim * d1d ** dim)
    mk = _i32(markers) if markers is not None else None
    lib().orc_domain_lf(dim, d1d, q1d, vdim, ne, _p(mk) if mk is not None else None, _p(_f64(B)),
                        _p(_f64(W)), _p(_f64(detJ)), _p(coeff), int(coeff.size == vdim), _p(y))
    return y


def randomize(n, seed):
    v = np.empty(n)
    lib().orc_randomize(n, seed, _p(v))
    return v


class Operator:
    """y = R^T A_E R x with the oracle kernels (+ ConstrainedOperator wrapper)."""

    def __init__(self, dim, d1d, q1d, ne, ndofs, gather_map, offsets, indices, B, G, D,
                 is_mass=False, ess=None, symmetric=1):
        self.keep = dict(g=_i32(gather_map), o=_i32(offsets), i=_i32(indices), B=_f64(B),
                         G=_f64(G if G is not None else B), D=_f64(D),
                         ess=_i32(ess if ess is not None else []))
        dof = d1d ** dim
        self.keep["xe"] = np.empty(ne * dof)
        self.keep["ye"] = np.empty(ne * dof)
        self.keep["z"] = np.empty(ndofs)
        k = self.keep
        self.ndofs = ndofs
        self.op = OrcOperator(dim, d1d, q1d, int(is_mass), symmetric, ne, ndofs,
                              k["g"].ctypes.data, k["o"].ctypes.data, k["i"].ctypes.data,
                              k["B"].ctypes.data, k["G"].ctypes.data, k["D"].ctypes.data,
                              k["ess"].ctypes.data, len(k["ess"]),
                              k["xe"].ctypes.data, k["ye"].ctypes.data, k["z"].ctypes.data)

    def mult(self, x):
        y = np.empty(self.ndofs)
        lib().orc_op_mult(C.byref(self.op), _p(_f64(x)), _p(y))
        return y

    def mult_constrained(self, x):
        y = np.empty(self.ndofs)
        lib().orc_op_mult_constrained(C.byref(self.op), _p(_f64(x)), _p(y))
        return y

    def diag(self):
        d = np.empty(self.ndofs)
        lib().orc_op_diag(C.byref(self.op), _p(d))
        return d

    def eliminate_rhs(self, x, b):
        b = _f64(b).copy()
        lib().orc_op_eliminate_rhs(C.byref(self.op), _p(_f64(x)), _p(b))
        return b

    def jacobi(self, diag, damping=1.0):
        dinv = np.empty(self.ndofs)
        e = self.keep["ess"]
        lib().orc_jacobi_setup(self.ndofs, _p(_f64(diag)), _p(e), len(e),
                               C.c_double(damping), _p(dinv))
        return dinv

    def pcg(self, b, dinv=None, rel_tol=1e-8, abs_tol=0.0, max_iter=2000):
        x = np.zeros(self.ndofs)
        hist = np.zeros(max_iter + 2)
        conv = C.c_int(0)
        fn = C.c_double(0)
        f = lib().orc_pcg
        f.restype = C.c_int
        its = f(C.byref(self.op), _p(_f64(dinv)) if dinv is not None else None, _p(_f64(b)),
                _p(x), C.c_double(rel_tol), C.c_double(abs_tol), max_iter, C.byref(conv),
                C.byref(fn), _p(hist), len(hist))
        return x, its, bool(conv.value), fn.value, hist[:its + 1]

    # OperatorChebyshevSmoother (linalg/solvers.cpp:453-658)
    def chebyshev_mult(self, dinv, order, max_eig, x):
        c = chebyshev_coeffs(order, max_eig)
        y = np.empty(self.ndofs)
        lib().orc_chebyshev_mult(C.byref(self.op), _p(_f64(dinv)), order, _p(c), _p(_f64(x)), _p(y))
        return y

    def power_method(self, dinv, start, steps=10, tol=1e-8):
        v = _f64(start).copy()
        f = lib().orc_power_method
        f.restype = C.c_double
        lam = f(C.byref(self.op), _p(_f64(dinv)), _p(v), steps, C.c_double(tol))
        return lam, v

    def pcg_chebyshev(self, b, dinv, order, max_eig, rel_tol=1e-8, abs_tol=0.0, max_iter=2000):
        c = chebyshev_coeffs(order, max_eig)
        x = np.zeros(self.ndofs)
        hist = np.zeros(max_iter + 2)
        conv = C.c_int(0)
        fn = C.c_double(0)
        f = lib().orc_pcg_chebyshev
        f.restype = C.c_int
        its = f(C.byref(self.op), _p(_f64(dinv)), order, _p(c), _p(_f64(b)), _p(x), C.c_double(rel_tol),
                C.c_double(abs_tol), max_iter, C.byref(conv), C.byref(fn), _p(hist), len(hist))
        return x, its, bool(conv.value), fn.value, hist[:its + 1]


def abs_mult(dim, d1d, q1d, ne, ndofs, gather_map, offsets, indices, B, G, D, x, is_mass=False, ess=None):
    """Operator::AbsMult of the PA operator (PABilinearFormExtension::MultInternal with useAbs,
    fem/bilinearform_ext.cpp:487-564; AddAbsMultPA: |B|, |G|, |pa_data|), optionally behind
    ConstrainedOperator::AbsMult (linalg/operator.cpp:663-722).  Returns (y, y_e)."""
    dof = d1d ** dim
    x = _f64(x)
    z = x.copy()
    if ess is not None and len(ess):
        z[ess] = 0.0
    g = _i32(gather_map)
    xe = z[np.where(g >= 0, g, -1 - g)]                      # |R|: signs dropped (restriction.cpp:138-157)
    Ba, Ga, Da = np.abs(_f64(B)), np.abs(_f64(G)), np.abs(_f64(D))
    if is_mass:
        ye = mass_apply(dim, d1d, q1d, ne, Ba, Da, xe)
    else:
        ye = diffusion_apply(dim, d1d, q1d, ne, Ba, Ga, Da, xe)
    y = restriction_abs_mult_transpose(ndofs, dof, offsets, indices, ye)
    if ess is not None and len(ess):
        y[ess] = x[ess]
    return y, ye


class MarkedForm:
    """BilinearForm with PA domain integrators restricted by element attribute
    (PABilinearFormExtension::MultInternal / AddMultWithMarkers / AssembleDiagonal,
    fem/bilinearform_ext.cpp:370-454,487-564,753-847), on top of the oracle's element kernels.
    integrators: list of (kind, D, marker-or-None), kind in {"diffusion", "mass"}."""

    def __init__(self, dim, d1d, q1d, ne, ndofs, gather_map, offsets, indices, B, G, attributes, integrators):
        self.dim, self.d1d, self.q1d, self.ne, self.ndofs = dim, d1d, q1d, ne, ndofs
        self.g, self.o, self.i = _i32(gather_map), _i32(offsets), _i32(indices)
        self.B, self.G = _f64(B), _f64(G)
        self.attr = _i32(attributes)
        self.integ = integrators
        self.dof = d1d ** dim

    def _on(self, marker):
        a = self.attr
        m = _i32(marker)
        return (a > 0) & (m[np.maximum(a, 1) - 1] != 0)        # bilinearform_ext.cpp:768

    def _apply(self, kind, D, xe):
        if kind == "mass":
            return mass_apply(self.dim, self.d1d, self.q1d, self.ne, self.B, D, xe)
        return diffusion_apply(self.dim, self.d1d, self.q1d, self.ne, self.B, self.G, D, xe)

    def mult(self, x):
        xe = restriction_mult(self.ne, self.dof, self.g, x)
        ye = np.zeros_like(xe)
        for kind, D, marker in self.integ:
            t = self._apply(kind, D, xe)                        # into a zeroed temporary (:818-829)
            if marker is None:
                ye += t
            else:
                on = np.repeat(self._on(marker), self.dof)
                ye[on] += t[on]                                  # AddWithMarkers_ (:753-774)
        return restriction_mult_transpose(self.ndofs, self.dof, self.o, self.i, ye)

    def diag(self):
        de = np.zeros(self.ne * self.dof)
        for kind, D, marker in self.integ:
            if kind == "mass":
                de += mass_diag(self.dim, self.d1d, self.q1d, self.ne, self.B, D)
            else:
                de += diffusion_diag(self.dim, self.d1d, self.q1d, self.ne, self.B, self.G, D)
            if marker is not None:
                # the reference zeroes the ACCUMULATED E-diagonal on unmarked elements (:374-399)
                de[np.repeat(~self._on(marker), self.dof)] = 0.0
        return restriction_abs_mult_transpose(self.ndofs, self.dof, self.o, self.i, de)

    def mult_constrained(self, x, ess):
        z = _f64(x).copy()
        z[ess] = 0.0
        y = self.mult(z)
        y[ess] = x[ess]
        return y


def chebyshev_coeffs(order, max_eig):
    c = np.zeros(max(order, 1))
    if lib().orc_chebyshev_coeffs(order, C.c_double(max_eig), _p(c)) != 0:
        raise ValueError("Chebyshev smoother not implemented for order = %d" % order)
    return c


class VectorOperator:
    """Vector (vdim > 1) PA operator: CEED BP2/BP4/BP6.

    Restatement of the reference's vector path by composition of the pinned scalar pieces:
      * ElementRestriction::Mult / MultTranspose with vdim (fem/restriction.cpp:116-201):
        y_E(i, c, e) = x_L(t ? c : j, t ? j : c) with t = byVDIM, j = gather_map index --
        component c of the L-vector is x[c::vdim] (byVDIM) or x[c*ndofs:(c+1)*ndofs] (byNODES);
        the E-vector is laid out (dof, vdim, NE);
      * Vector{Diffusion,Mass}Integrator::AddMultPA (fem/integ/bilininteg_vecdiffusion_pa.cpp,
        bilininteg_vecmass_pa.cpp): the scalar PA kernel applied to each component with the
        same quadrature data;
      * ConstrainedOperator / OperatorJacobiSmoother / CGSolver as for scalars, on vdofs.
    """

    def __init__(self, scalar_op, vdim, byvdim, ess_vdofs=None):
        self.op, self.vdim, self.byvdim = scalar_op, vdim, bool(byvdim)
        self.nd = scalar_op.ndofs
        self.n = vdim * self.nd
        self.ess = _i32(ess_vdofs if ess_vdofs is not None else [])

    def comp(self, x, c):
        return _f64(x[c::self.vdim] if self.byvdim else x[c * self.nd:(c + 1) * self.nd])

    def put(self, y, c, v):
        if self.byvdim:
            y[c::self.vdim] = v
        else:
            y[c * self.nd:(c + 1) * self.nd] = v

    def to_evector(self, x, ne, dof, gather_map):
        """ElementRestriction::Mult with vdim: (dof, vdim, NE)"""
        out = np.empty((ne, self.vdim, dof))
        for c in range(self.vdim):
            out[:, c, :] = restriction_mult(ne, dof, gather_map, self.comp(x, c)).reshape(ne, dof)
        return out.ravel()

    def mult(self, x):
        y = np.empty(self.n)
        for c in range(self.vdim):
            self.put(y, c, self.op.mult(self.comp(x, c)))
        return y

    def mult_constrained(self, x):
        z = _f64(x).copy()
        z[self.ess] = 0.0
        y = self.mult(z)
        y[self.ess] = x[self.ess]
        return y

    def diag(self):
        d = self.op.diag()
        y = np.empty(self.n)
        for c in range(self.vdim):
            self.put(y, c, d)
        return y

    def pcg(self, b, dinv, rel_tol=1e-8, max_iter=2000):
        """linalg/solvers.cpp:869-1050 on vdofs (numpy vectors, oracle operator)"""
        x = np.zeros(self.n)
        r = _f64(b).copy()
        z = dinv * r
        d = z.copy()
        nom0 = nom = float(np.dot(d, r))
        r0 = max(nom * rel_tol * rel_tol, 0.0)
        if nom <= r0:
            return x, 0, True
        q = self.mult_constrained(d)
        den = float(np.dot(q, d))
        i = 1
        while True:
            alpha = nom / den
            x += alpha * d
            r -= alpha * q
            z = dinv * r
            betanom = float(np.dot(r, z))
            if betanom <= r0:
                return x, i, True
            i += 1
            if i > max_iter:
                return x, max_iter, False
            d = z + (betanom / nom) * d
            q = self.mult_constrained(d)
            den = float(np.dot(d, q))
            nom = betanom
