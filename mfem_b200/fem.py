"""mfem_b200/fem.py -- Python mirror of the reference's host-side interface for the PA path.

Same names and call sequence as an MFEM driver (examples/ex1.cpp:122-267,
tests/benchmarks/bench_ceed.cpp:100-256): Device, Mesh.MakeCartesian3D, H1_FECollection,
FiniteElementSpace, IntRules, DiffusionIntegrator / MassIntegrator (AssemblePA, AddMultPA,
AssembleDiagonalPA), BilinearForm (SetAssemblyLevel, Assemble, Mult, AssembleDiagonal,
FormLinearSystem), ConstrainedOperator, OperatorJacobiSmoother, CGSolver.  Everything that
computes goes through the C ABI (include/b200fem.h) into the CUDA library; this file only
owns handles and torch device tensors (torch = device memory + streams, nothing else).
The C++ twin of this layer, compiled against MFEM's own headers, is adapter/.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import call, vp

KIND_DIFFUSION, KIND_MASS = 0, 1


def _np_ptr(a):
    return a.ctypes.data_as(vp)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---- host-only tables (no device needed) -------------------------------------------------
def gauss_legendre(n):
    p, w = np.empty(n), np.empty(n)
    call("b2f_gauss_legendre_h", n, _np_ptr(p), _np_ptr(w))
    return p, w


def gauss_lobatto(n):
    p, w = np.empty(n), np.empty(n)
    call("b2f_gauss_lobatto_h", n, _np_ptr(p), _np_ptr(w))
    return p, w


def h1_dof_to_quad(p, qpts):
    """DofToQuad TENSOR tables of the order-p H1 GLL basis: B, G of shape (nq*(p+1),), q fastest."""
    qpts = _f64(qpts)
    nq = len(qpts)
    B, G = np.empty(nq * (p + 1)), np.empty(nq * (p + 1))
    call("b2f_h1_dof_to_quad_h", p, nq, _np_ptr(qpts), _np_ptr(B), _np_ptr(G))
    return B, G


def tensor_weights(dim, w1):
    w1 = _f64(w1)
    W = np.empty(len(w1) ** dim)
    call("b2f_tensor_weights_h", dim, len(w1), _np_ptr(w1), _np_ptr(W))
    return W


class IntegrationRule:
    """Tensor-product rule on the unit square/cube (fem/intrules.cpp:43-96)."""

    def __init__(self, dim, pts1d, wts1d, order, gll=False):
        self.dim, self.pts1d, self.wts1d, self.order, self.gll = dim, pts1d, wts1d, order, gll
        self.nq1d = len(pts1d)
        self.weights = tensor_weights(dim, wts1d)

    def GetNPoints(self):
        return self.nq1d ** self.dim

    def GetWeights(self):
        return self.weights


class IntegrationRules:
    """IntRules.Get(geom, order) for SQUARE/CUBE (fem/intrules.cpp:1006-1090,1500-1560):
    Gauss-Legendre with n = order/2 + 1 points per direction; Gauss-Lobatto (the
    `IntegrationRules(0, Quadrature1D::GaussLobatto)` of bench_ceed.cpp:133) with
    n = (order+3)/2... i.e. the smallest n with 2n-3 >= order."""

    def __init__(self, gll=False):
        self.gll = gll

    def Get(self, dim, order):
        if self.gll:
            n = max(2, (order + 4) // 2)  # exact to degree 2n-3  (intrules.cpp:1023-1030)
            p, w = gauss_lobatto(n)
        else:
            n = order // 2 + 1
            p, w = gauss_legendre(n)
        return IntegrationRule(dim, p, w, order, self.gll)


IntRules = IntegrationRules(False)


# ---- device ------------------------------------------------------------------------------
class Device:
    """mfem::Device("cuda") (general/device.hpp:128): owns the library context.  No CPU
    backend exists: construction fails without a CUDA device."""

    def __init__(self, spec="cuda", device_id=0, stream=None):
        if not spec.startswith("cuda"):
            raise ValueError("mfem_b200 only has the 'cuda' backend (no CPU fallback)")
        import torch
        self.torch = torch
        if ":" in spec:
            device_id = int(spec.split(":")[1])
        self.device_id = device_id
        h = vp()
        call("b2f_ctx_create", device_id, vp(stream or 0), C.byref(h))
        self.ctx = h
        self.tdev = torch.device("cuda", device_id)
        torch.cuda.set_device(self.tdev)

    # memory helpers (torch tensors are the device allocations handed to the C ABI)
    def empty(self, n, dtype=None):
        return self.torch.empty(int(n), dtype=dtype or self.torch.float64, device=self.tdev)

    def zeros(self, n, dtype=None):
        return self.torch.zeros(int(n), dtype=dtype or self.torch.float64, device=self.tdev)

    def to_device(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.tdev)

    def sync(self):
        call("b2f_ctx_sync", self.ctx)

    def launch_count(self):
        return int(capi.lib().b2f_ctx_launch_count(self.ctx))

    def reset_launch_count(self):
        capi.lib().b2f_ctx_reset_launch_count(self.ctx)

    def sm_count(self):
        return int(capi.lib().b2f_ctx_sm_count(self.ctx))

    def close(self):
        if self.ctx:
            capi.lib().b2f_ctx_destroy(self.ctx)
            self.ctx = None


def _dp(t):
    return vp(t.data_ptr()) if t is not None else vp(0)


# ---- mesh + space ------------------------------------------------------------------------
def curve_map(X):
    """The smooth displacement of BASELINE config C5: x += 0.05 sin(2 pi y)
    sin(2 pi z), cyclically.  X is (..., dim)."""
    a, tp = 0.05, 2 * np.pi
    Y = X.copy()
    if X.shape[-1] == 3:
        Y[..., 0] += a * np.sin(tp * X[..., 1]) * np.sin(tp * X[..., 2])
        Y[..., 1] += a * np.sin(tp * X[..., 0]) * np.sin(tp * X[..., 2])
        Y[..., 2] += a * np.sin(tp * X[..., 0]) * np.sin(tp * X[..., 1])
    else:
        Y[..., 0] += a * np.sin(tp * X[..., 1])
        Y[..., 1] += a * np.sin(tp * X[..., 0])
    return Y


class Mesh:
    """Structured quad/hex mesh of the unit square/cube (Mesh::MakeCartesian2D/3D,
    mesh/mesh.cpp:4776,4786), optionally the block of one rank of a Cartesian partition."""

    def __init__(self, dim, nx, ny, nz=1, sfc_ordering=True, part=None, rank=0):
        self.dim, self.nx, self.ny, self.nz = dim, nx, ny, (nz if dim == 3 else 1)
        self.sfc = bool(sfc_ordering)
        self.part = tuple(part) if part else None  # (px, py, pz)
        self.rank = rank
        self.node_order = 1
        self.transform = None

    @staticmethod
    def MakeCartesian3D(nx, ny, nz, sfc_ordering=True):
        return Mesh(3, nx, ny, nz, sfc_ordering)

    @staticmethod
    def MakeCartesian2D(nx, ny, sfc_ordering=True):
        return Mesh(2, nx, ny, 1, sfc_ordering)

    def Dimension(self):
        return self.dim

    def SetCurvature(self, order):
        self.node_order = order

    def Transform(self, f):
        self.transform = f

    def nodes_e(self, elem_vertices):
        """Nodal coordinates as an E-vector (nd^dim, dim, NE), lexicographic -- what
        ElementRestriction::Mult of the nodes GridFunction gives (mesh/mesh.cpp:15445)."""
        dim, nv = self.dim, 2 ** self.dim
        ne = elem_vertices.size // (nv * dim)
        ev = elem_vertices.reshape(ne, dim, nv)
        if self.node_order == 1 and self.transform is None:
            return elem_vertices, 2
        nd = self.node_order + 1
        g, _ = gauss_lobatto(nd)
        lo, hi = ev[:, :, 0], ev[:, :, nv - 1]  # opposite corners (axis-aligned boxes)
        axes = [lo[:, a, None] + (hi[:, a] - lo[:, a])[:, None] * g[None, :] for a in range(dim)]
        if dim == 3:
            X = np.stack(np.broadcast_arrays(axes[0][:, None, None, :], axes[1][:, None, :, None],
                                             axes[2][:, :, None, None]), axis=-1)
        else:
            X = np.stack(np.broadcast_arrays(axes[0][:, None, :], axes[1][:, :, None]), axis=-1)
        if self.transform is not None:
            X = self.transform(X)
        # (ne, [z,] y, x, dim) -> (ne, dim, nd^dim)
        X = np.moveaxis(X.reshape(ne, nd ** dim, dim), 2, 1)
        return np.ascontiguousarray(X).ravel(), nd


class H1_FECollection:
    def __init__(self, order, dim):
        self.order, self.dim = order, dim


class FiniteElementSpace:
    """Scalar H1 space on a structured Mesh with the reference's dof numbering
    (fem/fespace.cpp:3428-3536), built by the library's host code."""

    def __init__(self, mesh, fec, vdim=1, ordering="byNODES"):
        """vdim > 1: vector space (FiniteElementSpace(mesh, fec, vdim, ordering), fem/fespace.hpp);
        bench_ceed.cpp:131 uses Ordering::byVDIM for the vector BPs."""
        self.mesh, self.fec = mesh, fec
        self.vdim, self.byvdim = int(vdim), ordering == "byVDIM"
        if vdim > 1 and mesh.part is not None and not self.byvdim:
            raise ValueError("rank-parallel vector spaces need Ordering::byVDIM")
        p, dim = fec.order, mesh.dim
        self.p, self.dim = p, dim
        self.d1d = p + 1
        self.dof = self.d1d ** dim
        nv = 2 ** dim
        self.halo_tables = None
        if mesh.part is None:
            ne, nd, nb = C.c_int(), C.c_int(), C.c_int()
            call("b2f_cartesian_h1_space_h", dim, mesh.nx, mesh.ny, mesh.nz, p, int(mesh.sfc),
                 C.byref(ne), C.byref(nd), None, None, None, None, None)
            self.ne, self.ndofs = ne.value, nd.value
            self.elem_dof = np.empty(self.ne * self.dof, np.int32)
            self.dof_map = np.empty(self.dof, np.int32)
            self.elem_vertices = np.empty(self.ne * nv * dim)
            call("b2f_cartesian_h1_space_h", dim, mesh.nx, mesh.ny, mesh.nz, p, int(mesh.sfc),
                 C.byref(ne), C.byref(nd), _np_ptr(self.elem_dof), _np_ptr(self.dof_map),
                 C.byref(nb), None, _np_ptr(self.elem_vertices))
            self.bdr_dofs = np.empty(nb.value, np.int32)
            call("b2f_cartesian_h1_space_h", dim, mesh.nx, mesh.ny, mesh.nz, p, int(mesh.sfc),
                 C.byref(ne), C.byref(nd), None, None, C.byref(nb), _np_ptr(self.bdr_dofs), None)
            self.ntdofs = self.ndofs
        else:
            px, py, pz = mesh.part
            sz = (C.c_int * 11)()
            args = (dim, mesh.nx, mesh.ny, mesh.nz, p, int(mesh.sfc), px, py, pz, mesh.rank)
            call("b2f_cartesian_partition_h", *args, sz, *([None] * 11))
            lnx, lny, lnz, ne, nl, nt, nn, ns, nx_, nb, self.n_interior = list(sz)
            self.local_n = (lnx, lny, lnz)
            self.ne, self.ndofs, self.ntdofs = ne, nl, nt
            self.elem_dof = np.empty(ne * self.dof, np.int32)
            self.dof_map = np.empty(self.dof, np.int32)
            self.bdr_dofs = np.empty(nb, np.int32)
            self.elem_vertices = np.empty(ne * nv * dim)
            t = dict(ltdof_ldof=np.empty(nt, np.int32), nbr_rank=np.empty(nn, np.int32),
                     shr_offsets=np.empty(nn + 1, np.int32), shr_ltdof=np.empty(ns, np.int32),
                     ext_offsets=np.empty(nn + 1, np.int32), ext_ldof=np.empty(nx_, np.int32),
                     ldof_lattice=np.empty(nl, np.int64))
            call("b2f_cartesian_partition_h", *args, sz, _np_ptr(self.elem_dof),
                 _np_ptr(self.dof_map), _np_ptr(self.bdr_dofs), _np_ptr(self.elem_vertices),
                 _np_ptr(t["ltdof_ldof"]), _np_ptr(t["nbr_rank"]), _np_ptr(t["shr_offsets"]),
                 _np_ptr(t["shr_ltdof"]), _np_ptr(t["ext_offsets"]), _np_ptr(t["ext_ldof"]),
                 _np_ptr(t["ldof_lattice"]))
            self.halo_tables = t
        self._restr = None

    def GetNE(self):
        return self.ne

    def GetNDofs(self):
        return self.ndofs

    def GetTrueVSize(self):
        return self.ntdofs * self.vdim

    def GetVSize(self):
        return self.ndofs * self.vdim

    def GetEssentialTrueDofs(self):
        """All-boundary essential dofs (ex1.cpp:165-183 with ess_bdr = 1): LOCAL (v)dof ids."""
        if self.vdim == 1:
            return self.bdr_dofs
        b = self.bdr_dofs.astype(np.int64)
        if self.byvdim:
            v = (b[:, None] * self.vdim + np.arange(self.vdim)[None, :]).ravel()
        else:
            v = (b[None, :] + (np.arange(self.vdim) * self.ndofs)[:, None]).ravel()
        return np.sort(v).astype(np.int32)

    def GetElementRestriction(self, device):
        if self._restr is None:
            self._restr = ElementRestriction.from_space(device, self)
        return self._restr


class ElementRestriction:
    """fem/restriction.hpp:40 -- L-vector <-> E-vector."""

    def __init__(self, device, handle, ne, dof, ndofs, vdim=1):
        self.dev, self.h, self.ne, self.dof, self.ndofs, self.vdim = device, handle, ne, dof, ndofs, vdim

    @classmethod
    def from_space(cls, device, fes):
        h = vp()
        vdim, byvdim = getattr(fes, "vdim", 1), int(getattr(fes, "byvdim", False))
        call("b2f_restriction_create_from_elem_dof", device.ctx, fes.ne, fes.dof, fes.ndofs, vdim, byvdim,
             _np_ptr(fes.elem_dof), _np_ptr(fes.dof_map), C.byref(h))
        return cls(device, h, fes.ne, fes.dof, fes.ndofs, vdim)

    @classmethod
    def from_tables(cls, device, ne, dof, ndofs, gather_map, offsets, indices, vdim=1, byvdim=False):
        h = vp()
        g, o, i = _i32(gather_map), _i32(offsets), _i32(indices)
        call("b2f_restriction_create", device.ctx, ne, dof, ndofs, vdim, int(bool(byvdim)), _np_ptr(g),
             _np_ptr(o), _np_ptr(i), C.byref(h))
        return cls(device, h, ne, dof, ndofs, vdim)

    def tables(self):
        """(gather_map, offsets, indices) copied back to the host."""
        import torch
        g, o, i = vp(), vp(), vp()
        call("b2f_restriction_tables", self.h, C.byref(g), C.byref(o), C.byref(i))
        out = []
        for ptr, n in ((g, self.ne * self.dof), (o, self.ndofs + 1), (i, self.ne * self.dof)):
            a = np.empty(n, np.int32)
            call("b2f_memcpy_d2h", self.dev.ctx, _np_ptr(a), ptr, a.nbytes)
            out.append(a)
        self.dev.sync()
        return tuple(out)

    def Height(self):
        return self.ne * self.dof * self.vdim

    def Mult(self, x, y):
        call("b2f_restriction_mult", self.h, _dp(x), _dp(y))

    def AbsMult(self, x, y):
        call("b2f_restriction_abs_mult", self.h, _dp(x), _dp(y))

    def MultTranspose(self, x, y):
        call("b2f_restriction_mult_transpose", self.h, _dp(x), _dp(y), 0)

    def AddMultTranspose(self, x, y):
        call("b2f_restriction_mult_transpose", self.h, _dp(x), _dp(y), 1)

    def AbsMultTranspose(self, x, y):
        call("b2f_restriction_abs_mult_transpose", self.h, _dp(x), _dp(y))

    def close(self):
        if self.h:
            capi.lib().b2f_restriction_destroy(self.h)
            self.h = None


# ---- integrators -------------------------------------------------------------------------
class _PAIntegrator:
    kind = None

    def __init__(self, coeff=1.0, ir=None):
        self.coeff, self.ir, self.h, self.dev = float(coeff), ir, None, None

    def GetRule(self, fes):
        # DiffusionIntegrator::GetRule bilininteg.cpp:1347 (2p + dim - 1) and
        # MassIntegrator::GetRule :1459 (2p + dim - 1 on affine tensor meshes): Q = p + 2 in 3D
        return IntRules.Get(fes.dim, 2 * fes.p + fes.dim - 1)

    def AssemblePA(self, device, fes, weights=None):
        """AssemblePA (bilininteg_diffusion_pa.cpp:108 / bilininteg_mass_pa.cpp:25) with the
        geometric factors fused in: D is computed straight from the nodes, J never exists.
        weights: quadrature weights to use instead of the rule's (LinearForm passes ones to
        obtain detJ itself)."""
        self.dev = device
        ir = self.ir or self.GetRule(fes)
        self.ir = ir
        self.d1d, self.q1d, self.dim, self.ne = fes.d1d, ir.nq1d, fes.dim, fes.ne
        self.B, self.G = h1_dof_to_quad(fes.p, ir.pts1d)
        nodes, nd = fes.mesh.nodes_e(fes.elem_vertices)
        Bn, Gn = h1_dof_to_quad(nd - 1, ir.pts1d)
        nodes_d = device.to_device(_f64(nodes))
        h = vp()
        call("b2f_pa_setup_from_nodes", device.ctx, self.kind, fes.dim, fes.d1d, ir.nq1d, fes.ne,
             _np_ptr(self.B), _np_ptr(self.G), _np_ptr(_f64(ir.weights if weights is None else weights)), nd,
             _np_ptr(Bn), _np_ptr(Gn), _dp(nodes_d), self.coeff, C.byref(h))
        self.h = h
        device.sync()
        del nodes_d

    def assemble_from_factors(self, device, dim, d1d, q1d, ne, B, G, W, J=None, detJ=None,
                              coeff=None, coeff_dim=1):
        """The reference seam itself: D from GeometricFactors J / detJ
        (PADiffusionSetup, bilininteg_diffusion_kernels.cpp:107; mass_pa.cpp:73)."""
        self.dev = device
        self.d1d, self.q1d, self.dim, self.ne = d1d, q1d, dim, ne
        self.B, self.G = _f64(B), _f64(G if G is not None else B)
        W = _f64(W)
        c = _f64(coeff if coeff is not None else [self.coeff])
        c_d = device.to_device(c)
        h = vp()
        if self.kind == KIND_DIFFUSION:
            J_d = device.to_device(_f64(J))
            call("b2f_diffusion_setup", device.ctx, dim, d1d, q1d, ne, _np_ptr(self.B), _np_ptr(self.G),
                 _np_ptr(W), _dp(J_d), _dp(c_d), coeff_dim, c.size, C.byref(h))
        else:
            d_d = device.to_device(_f64(detJ))
            call("b2f_mass_setup", device.ctx, dim, d1d, q1d, ne, _np_ptr(self.B), _np_ptr(W),
                 _dp(d_d), _dp(c_d), c.size, C.byref(h))
        self.h = h
        device.sync()

    def pa_data(self):
        ptr, n = vp(), capi.i64()
        call("b2f_pa_data", self.h, C.byref(ptr), C.byref(n))
        a = np.empty(n.value)
        call("b2f_memcpy_d2h", self.dev.ctx, _np_ptr(a), ptr, a.nbytes)
        self.dev.sync()
        return a

    def AddMultPA(self, x_e, y_e, vdim=1):
        """vdim > 1: Vector{Diffusion,Mass}Integrator::AddMultPA on (dofs, vdim, NE) E-vectors"""
        if vdim == 1:
            call("b2f_pa_add_mult", self.h, _dp(x_e), _dp(y_e))
        else:
            call("b2f_pa_add_mult_vec", self.h, vdim, _dp(x_e), _dp(y_e))

    def AddAbsMultPA(self, x_e, y_e):
        call("b2f_pa_add_abs_mult", self.h, _dp(x_e), _dp(y_e))

    def AddMultTransposePA(self, x_e, y_e):
        call("b2f_pa_add_mult_transpose", self.h, _dp(x_e), _dp(y_e))

    def AssembleDiagonalPA(self, diag_e):
        call("b2f_pa_assemble_diagonal", self.h, _dp(diag_e))

    def close(self):
        if self.h:
            capi.lib().b2f_pa_destroy(self.h)
            self.h = None


class DiffusionIntegrator(_PAIntegrator):
    kind = KIND_DIFFUSION


class MassIntegrator(_PAIntegrator):
    kind = KIND_MASS

    def GetRule(self, fes):
        return IntRules.Get(fes.dim, 2 * fes.p + fes.dim - 1)


# Vector{Diffusion,Mass}Integrator (fem/integ/bilininteg_vecdiffusion_pa.cpp, bilininteg_vecmass_pa.cpp):
# the same quadrature data; the operator applies it to every component of a vector space.
VectorDiffusionIntegrator = DiffusionIntegrator
VectorMassIntegrator = MassIntegrator


# ---- rank-parallel prolongation ----------------------------------------------------------
class Halo:
    """ConformingProlongationOperator (fem/pfespace.hpp:615) over NCCL."""

    @staticmethod
    def expand_vdim(tables, vdim):
        """Scalar shared-dof tables -> the vdof tables of a vdim-component space ordered byVDIM
        (vdof = ldof*vdim + c), which is what ParFiniteElementSpace's communicator holds for
        vdim > 1.  Owned-first numbering is preserved: true vdofs stay the head of the vector."""
        c = np.arange(vdim, dtype=np.int64)

        def ex(a):
            return _i32((np.asarray(a, np.int64)[:, None] * vdim + c[None, :]).ravel())
        t = dict(tables)
        t["ltdof_ldof"] = ex(tables["ltdof_ldof"])
        t["shr_ltdof"] = ex(tables["shr_ltdof"])
        t["ext_ldof"] = ex(tables["ext_ldof"])
        t["shr_offsets"] = _i32(np.asarray(tables["shr_offsets"], np.int64) * vdim)
        t["ext_offsets"] = _i32(np.asarray(tables["ext_offsets"], np.int64) * vdim)
        return t

    def __init__(self, device, rank, nranks, n_ldofs, tables, nccl_id=None, vdim=1):
        """vdim > 1: n_ldofs and tables are the SCALAR ones of the space; the halo is built on
        its vdofs (Ordering::byVDIM)."""
        t = tables if vdim == 1 else Halo.expand_vdim(tables, vdim)
        n_ldofs = n_ldofs * vdim
        self.vdim = vdim
        self.dev, self.tables = device, t
        self.n_ldofs, self.n_tdofs = n_ldofs, len(t["ltdof_ldof"])
        h = vp()
        idbuf = (C.c_char * 128).from_buffer_copy(nccl_id) if nccl_id is not None else None
        call("b2f_halo_create", device.ctx, idbuf, rank, nranks, n_ldofs, self.n_tdofs,
             _np_ptr(_i32(t["ltdof_ldof"])), len(t["nbr_rank"]), _np_ptr(_i32(t["nbr_rank"])),
             _np_ptr(_i32(t["shr_offsets"])), _np_ptr(_i32(t["shr_ltdof"])),
             _np_ptr(_i32(t["ext_offsets"])), _np_ptr(_i32(t["ext_ldof"])), C.byref(h))
        self.h = h

    def connect_p2p(self, dist):
        """Switch the exchange to peer-memory stores over NVLink (CUDA IPC): gathers every rank's
        buffer handles with torch.distributed and maps the neighbours' receive buffers."""
        nb = int(capi.lib().b2f_halo_p2p_blob_bytes())
        blob = (C.c_char * nb)()
        call("b2f_halo_p2p_export", self.h, blob)
        blobs = [None] * dist.get_world_size()
        dist.all_gather_object(blobs, bytes(blob))
        allb = (C.c_char * (nb * len(blobs))).from_buffer_copy(b"".join(blobs))
        call("b2f_halo_p2p_connect", self.h, allb, len(blobs))
        dist.barrier()

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        call("b2f_nccl_unique_id_h", buf)
        return bytes(buf)

    def Mult(self, x_t, y_l):
        call("b2f_halo_mult", self.h, _dp(x_t), _dp(y_l))

    def MultTranspose(self, x_l, y_t):
        call("b2f_halo_mult_transpose", self.h, _dp(x_l), _dp(y_t))

    def Dot(self, x, y):
        r = C.c_double()
        call("b2f_halo_dot", self.h, x.numel(), _dp(x), _dp(y), C.byref(r))
        return r.value

    def close(self):
        if self.h:
            capi.lib().b2f_halo_destroy(self.h)
            self.h = None


# ---- bilinear form -----------------------------------------------------------------------
class BilinearForm:
    """BilinearForm with AssemblyLevel::PARTIAL (fem/bilinearform.cpp:109-136 ->
    PABilinearFormExtension, fem/bilinearform_ext.cpp:256-564)."""

    def __init__(self, device, fes):
        self.dev, self.fes = device, fes
        self.integrators = []
        self.markers = []          # per integrator: None or the 0/1 array indexed by attribute - 1
        self.elem_attributes = None
        self.op = None
        self.ess = np.zeros(0, np.int32)
        self.halo = None

    def SetAssemblyLevel(self, level="PARTIAL"):
        if level != "PARTIAL":
            raise ValueError("only AssemblyLevel::PARTIAL exists on this path")

    def AddDomainIntegrator(self, integ, elem_marker=None):
        """elem_marker: BilinearForm::AddDomainIntegrator(bfi, elem_marker) -- the integrator acts
        on the elements whose attribute a has elem_marker[a-1] != 0; the element attributes
        come from SetElementAttributes (Mesh::GetElementAttributes, mesh/mesh.cpp:1040)."""
        self.integrators.append(integ)
        self.markers.append(None if elem_marker is None else _i32(elem_marker))

    def SetElementAttributes(self, attributes):
        self.elem_attributes = _i32(attributes)

    def Assemble(self):
        self.restr = self.fes.GetElementRestriction(self.dev)
        for it in self.integrators:
            if it.h is None:
                it.AssemblePA(self.dev, self.fes)
        self._make_op(self.ess)

    def _make_op(self, ess):
        if self.op:
            capi.lib().b2f_operator_destroy(self.op)
            self.op = None
        ess = _i32(ess)
        arr = (vp * len(self.integrators))(*[it.h for it in self.integrators])
        h = vp()
        call("b2f_operator_create", self.dev.ctx, self.restr.h, arr, len(self.integrators),
             _np_ptr(ess) if len(ess) else None, len(ess), C.byref(h))
        self.op, self.ess = h, ess
        for i, mk in enumerate(self.markers):
            if mk is not None:
                if self.elem_attributes is None:
                    raise ValueError("AddDomainIntegrator(elem_marker): call SetElementAttributes first")
                call("b2f_operator_set_markers", self.op, i, _np_ptr(self.elem_attributes), _np_ptr(mk), len(mk))
        if self.halo is not None:
            call("b2f_operator_set_halo", self.op, self.halo.h)
            self._set_interior()

    def SetHalo(self, halo):
        """ParBilinearForm: the operator becomes P^T A P on true-dof vectors."""
        self.halo = halo
        if self.op:
            call("b2f_operator_set_halo", self.op, halo.h)
            self._set_interior()

    def _set_interior(self):
        ni = getattr(self.fes, "n_interior", None)
        if ni is not None and self.halo is not None:
            call("b2f_operator_set_interior", self.op, ni)

    def Height(self):
        return int(capi.lib().b2f_operator_height(self.op))

    def constrained_tdofs(self, ess_dofs=None):
        """The operator's own constrained list: true dofs when a halo is attached."""
        ess = self.ess if ess_dofs is None else ess_dofs
        if self.halo is None:
            return _i32(ess)
        l2t = np.full(self.halo.n_ldofs, -1, np.int64)
        l2t[self.halo.tables["ltdof_ldof"]] = np.arange(self.halo.n_tdofs)
        e = l2t[np.asarray(ess, np.int64)]
        return _i32(e[e >= 0])

    def Mult(self, x, y):
        call("b2f_operator_mult", self.op, _dp(x), _dp(y))

    def MultUnfused(self, x, y):
        call("b2f_operator_mult_unfused", self.op, _dp(x), _dp(y))

    def AbsMult(self, x, y):
        call("b2f_operator_abs_mult", self.op, _dp(x), _dp(y))

    def AssembleDiagonal(self, diag):
        call("b2f_operator_assemble_diagonal", self.op, _dp(diag))

    def FormLinearSystem(self, ess_ldofs, x, b):
        """Operator::FormLinearSystem (linalg/operator.cpp:130): returns the
        ConstrainedOperator A; b is modified in place into B (EliminateRHS), X aliases x."""
        self._make_op(ess_ldofs)
        A = ConstrainedOperator(self)
        A.EliminateRHS(x, b)
        return A, x, b

    def close(self):
        if self.op:
            capi.lib().b2f_operator_destroy(self.op)
            self.op = None


class ConstrainedOperator:
    """linalg/operator.cpp:526-660; diag_policy: "DIAG_ONE" (default), "DIAG_ZERO", "DIAG_KEEP"."""

    def __init__(self, form, diag_policy="DIAG_ONE"):
        self.form = form
        call("b2f_operator_set_diag_policy", form.op, {"DIAG_ZERO": 0, "DIAG_ONE": 1, "DIAG_KEEP": 2}[diag_policy])

    def AssembleDiagonal(self, diag):
        call("b2f_operator_assemble_diagonal_constrained", self.form.op, _dp(diag))

    def Height(self):
        return self.form.Height()

    def Mult(self, x, y):
        call("b2f_operator_mult_constrained", self.form.op, _dp(x), _dp(y))

    def AbsMult(self, x, y):
        call("b2f_operator_abs_mult_constrained", self.form.op, _dp(x), _dp(y))

    def EliminateRHS(self, x, b):
        call("b2f_operator_eliminate_rhs", self.form.op, _dp(x), _dp(b))


class OperatorJacobiSmoother:
    """linalg/solvers.cpp:324-450."""

    def __init__(self, form, ess_dofs=None, damping=1.0):
        dev = form.dev
        n = form.Height()
        self.dev, self.n = dev, n
        diag = dev.empty(n)
        form.AssembleDiagonal(diag)
        self.dinv = dev.empty(n)
        ess = form.constrained_tdofs(ess_dofs)
        ess_d = dev.to_device(ess) if len(ess) else None
        call("b2f_jacobi_setup", dev.ctx, n, _dp(diag), _dp(ess_d), len(ess), float(damping),
             _dp(self.dinv))
        dev.sync()

    def Mult(self, x, y):
        call("b2f_jacobi_mult", self.dev.ctx, self.n, _dp(self.dinv), _dp(x), _dp(y))


class OperatorChebyshevSmoother:
    """linalg/solvers.hpp:392-505, solvers.cpp:453-658.  `oper` is the ConstrainedOperator of a
    BilinearForm; `diag` its assembled diagonal (device vector).  Either give
    `max_eig_estimate`, or leave it None to run the power method from `start` (a host vector
    the caller draws: Vector::Randomize(power_seed) in the reference)."""

    def __init__(self, oper, diag, order, max_eig_estimate=None, power_iterations=10,
                 power_tolerance=1e-8, start=None):
        form = oper.form if hasattr(oper, "form") else oper
        dev = form.dev
        self.dev, self.form, self.order = dev, form, order
        self.n = form.Height()
        if max_eig_estimate is None:
            if start is None:
                raise ValueError("OperatorChebyshevSmoother: the power method needs a start vector")
            jac = OperatorJacobiSmoother.__new__(OperatorJacobiSmoother)
            jac.dev, jac.n = dev, self.n
            jac.dinv = dev.empty(self.n)
            ess = form.constrained_tdofs()
            ess_d = dev.to_device(ess) if len(ess) else None
            call("b2f_jacobi_setup", dev.ctx, self.n, _dp(diag), _dp(ess_d), len(ess), 1.0, _dp(jac.dinv))
            v0 = dev.to_device(np.ascontiguousarray(start, dtype=np.float64))
            lam = C.c_double()
            call("b2f_power_method", form.op, _dp(jac.dinv), _dp(v0), int(power_iterations),
                 float(power_tolerance), C.byref(lam))
            max_eig_estimate = lam.value
        self.max_eig_estimate = float(max_eig_estimate)
        h = C.c_void_p()
        call("b2f_chebyshev_create", form.op, _dp(diag), int(order), self.max_eig_estimate, C.byref(h))
        self.handle = h
        dev.sync()

    @staticmethod
    def coefficients(order, max_eig_estimate):
        c = np.zeros(5)
        call("b2f_chebyshev_coefficients_h", int(order), float(max_eig_estimate), _np_ptr(c))
        return c[:order]

    def Mult(self, x, y):
        call("b2f_chebyshev_mult", self.handle, _dp(x), _dp(y))

    def close(self):
        if self.handle:
            capi.lib().b2f_chebyshev_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LORSparseMatrix:
    """The assembled LOR matrix (mfem::SparseMatrix, CSR) resident on the device."""

    def __init__(self, device, handle):
        self.dev, self.h = device, handle
        n, nnz = C.c_int(), capi.i64()
        self._I, self._J, self._A = vp(), vp(), vp()
        call("b2f_lor_csr", handle, C.byref(n), C.byref(nnz), C.byref(self._I), C.byref(self._J), C.byref(self._A))
        self.n, self.nnz = n.value, nnz.value

    def Height(self):
        return self.n

    def NumNonZeroElems(self):
        return self.nnz

    def kernel_ms(self):
        """device time of the kernels of the assembly (allocations excluded)"""
        ms = C.c_double()
        call("b2f_lor_kernel_ms", self.h, C.byref(ms))
        return ms.value

    def device_pointers(self):
        """(I, J, data) device addresses, e.g. for a multigrid set-up that takes CSR in place"""
        return self._I.value, self._J.value, self._A.value

    def _d2h(self, ptr, n, dtype):
        a = np.empty(n, dtype=dtype)
        if n:
            call("b2f_memcpy_d2h", self.dev.ctx, _np_ptr(a), ptr, a.nbytes)
            self.dev.sync()
        return a

    def GetI(self):
        return self._d2h(self._I, self.n + 1, np.int32)

    def GetJ(self):
        return self._d2h(self._J, self.nnz, np.int32)

    def GetData(self):
        return self._d2h(self._A, self.nnz, np.float64)

    def sparse_ij(self):
        """BatchedLORAssembly::sparse_ij, (NE, (p+1)^dim, 3^dim); needs keep_sparse_ij at assembly"""
        ptr, n = vp(), capi.i64()
        call("b2f_lor_sparse_ij", self.h, C.byref(ptr), C.byref(n))
        return self._d2h(ptr, n.value, np.float64)

    def Mult(self, x, y):
        call("b2f_lor_mult", self.h, _dp(x), _dp(y))

    def close(self):
        if self.h:
            capi.lib().b2f_lor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchedLORAssembly:
    """BatchedLORAssembly (fem/lor/lor_batched.hpp:26-131) for H1 spaces in 2D / 3D: the low-order-refined
    matrix of a mass + diffusion form, assembled on the device as CSR -- what LORSolver hands to an
    algebraic multigrid.  The constructor forms the LOR vertex coordinates (FormLORVertexCoordinates)."""

    def __init__(self, device, fes):
        if fes.dim not in (2, 3) or fes.vdim != 1:
            raise ValueError("BatchedLORAssembly: scalar H1 spaces in 2D / 3D only")
        self.dev, self.fes = device, fes
        p, ne = fes.p, fes.ne
        nodes, nd = fes.mesh.nodes_e(fes.elem_vertices)
        gp, _ = gauss_lobatto(p + 1)
        Bn, _ = h1_dof_to_quad(nd - 1, gp)  # (q, d) column-major: B[q + nq*d]
        Bn = np.ascontiguousarray(_f64(Bn).reshape(nd, p + 1).T)  # (p+1, nd) row-major
        nodes_d = device.to_device(_f64(nodes))
        self.x_vert = device.empty(fes.dim * (p + 1) ** fes.dim * ne)
        call("b2f_lor_vertex_coordinates", device.ctx, fes.dim, p, nd, ne, _np_ptr(Bn), _dp(nodes_d),
             _dp(self.x_vert))
        del nodes_d

    def GetLORVertexCoordinates(self):
        return self.x_vert

    def Assemble(self, a, ess_dofs=None, mass_coeff=None, diff_coeff=None, keep_sparse_ij=False):
        """Assemble(a, ess_dofs, A): `a` is a BilinearForm whose MassIntegrator / DiffusionIntegrator
        constants are used, or None with mass_coeff / diff_coeff given: a number, or the coefficient at
        the (p+1)^dim Gauss-Lobatto points of every element, (NE, (p+1)^dim) -- BatchedLORKernel's c1 / c2."""
        dev, fes = self.dev, self.fes
        cm, cd = 0.0, 0.0
        if a is not None:
            for it in a.integrators:
                if isinstance(it, MassIntegrator):
                    cm = it.coeff
                elif isinstance(it, DiffusionIntegrator):
                    cd = it.coeff
                else:
                    raise ValueError("BatchedLORAssembly: MassIntegrator / DiffusionIntegrator only")
        cm = cm if mass_coeff is None else mass_coeff
        cd = cd if diff_coeff is None else diff_coeff
        cm_d = dev.to_device(_f64(np.atleast_1d(cm)).ravel())
        cd_d = dev.to_device(_f64(np.atleast_1d(cd)).ravel())
        ess = _i32(ess_dofs if ess_dofs is not None else [])
        ess_d = dev.to_device(ess) if ess.size else None
        R = fes.GetElementRestriction(dev)
        h = vp()
        call("b2f_lor_assemble_h1", dev.ctx, R.h, fes.dim, fes.p, _dp(self.x_vert), _dp(cm_d), cm_d.numel(), _dp(cd_d),
             cd_d.numel(), _dp(ess_d), ess.size, int(bool(keep_sparse_ij)), C.byref(h))
        return LORSparseMatrix(dev, h)


class CGSolver:
    """linalg/solvers.cpp:869-1050 (iterative_mode = false)."""

    def __init__(self):
        self.rel_tol, self.abs_tol, self.max_iter = 0.0, 0.0, 10
        self.prec, self.oper = None, None
        self.check_every = 1
        self.history = None

    def SetRelTol(self, v):
        self.rel_tol = v

    def SetAbsTol(self, v):
        self.abs_tol = v

    def SetMaxIter(self, v):
        self.max_iter = v

    def SetOperator(self, A):
        self.oper = A

    def SetPreconditioner(self, M):
        self.prec = M

    def Mult(self, b, x, record_history=False):
        opt = capi.PcgOptions(self.rel_tol, self.abs_tol, self.max_iter, self.check_every)
        res = capi.PcgResult()
        form = self.oper.form
        hist = np.zeros(self.max_iter + 2) if record_history else None
        tail = (_dp(b), _dp(x), C.byref(opt), C.byref(res), _np_ptr(hist) if hist is not None else None,
                len(hist) if hist is not None else 0)
        if isinstance(self.prec, OperatorChebyshevSmoother):
            call("b2f_pcg_chebyshev", form.op, self.prec.handle, *tail)
        else:
            call("b2f_pcg", form.op, _dp(self.prec.dinv) if self.prec else None, *tail)
        self.final_iter, self.converged = res.iterations, bool(res.converged)
        self.initial_norm, self.final_norm = res.initial_norm, res.final_norm
        self.history = hist[:res.iterations + 1] if hist is not None else None

    def GetNumIterations(self):
        return self.final_iter

    def GetConverged(self):
        return self.converged

    def GetFinalNorm(self):
        return self.final_norm


def vector_dot(device, x, y):
    r = C.c_double()
    call("b2f_vector_dot", device.ctx, x.numel(), _dp(x), _dp(y), C.byref(r))
    return r.value


def domain_lf_assemble(device, dim, d1d, q1d, vdim, ne, B, W, detJ_d, coeff_d, y_e, markers_d=None):
    """DomainLFIntegrator / VectorDomainLFIntegrator::AssembleDevice (fem/integ/lininteg_domain.cpp:
    21-72) on device E-vectors: y_e += B^T (W f detJ)."""
    B, W = _f64(B), _f64(W)
    call("b2f_domain_lf_assemble", device.ctx, dim, d1d, q1d, vdim, ne, _dp(markers_d), _np_ptr(B), _np_ptr(W),
         _dp(detJ_d), _dp(coeff_d), coeff_d.numel(), _dp(y_e))


class LinearForm:
    """b_i = (1, phi_i): LinearForm + DomainLFIntegrator(one) (fem/lininteg.hpp:116, rule order
    2p + 0), assembled on the device as LinearForm::UseFastAssembly does
    (fem/linearform_ext.cpp:20-66): DomainLFIntegrator::AssembleDevice on E-vectors
    (b2f_domain_lf_assemble), then ElementRestriction::MultTranspose.  detJ at the rule's points
    comes from the fused geometric-factor kernel (mass set-up with unit weights: D = detJ)."""

    def __init__(self, device, fes):
        self.dev, self.fes = device, fes

    def Assemble(self):
        fes, dev = self.fes, self.dev
        ir = IntRules.Get(fes.dim, 2 * fes.p)
        g = MassIntegrator(1.0, ir)
        g.AssemblePA(dev, fes, weights=np.ones(ir.nq1d ** fes.dim))
        detJ, n = vp(), capi.i64()
        call("b2f_pa_data", g.h, C.byref(detJ), C.byref(n))
        R = fes.GetElementRestriction(dev)
        vd = getattr(fes, "vdim", 1)
        one = dev.empty(vd).fill_(1.0)
        ye = dev.zeros(R.Height())
        call("b2f_domain_lf_assemble", dev.ctx, fes.dim, fes.d1d, ir.nq1d, vd, fes.ne, vp(0), _np_ptr(_f64(g.B)),
             _np_ptr(_f64(ir.weights)), detJ, _dp(one), vd, _dp(ye))
        b = dev.empty(fes.ndofs * vd)
        R.MultTranspose(ye, b)
        dev.sync()
        g.close()
        return b
