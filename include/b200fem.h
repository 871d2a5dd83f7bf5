/* include/b200fem.h -- C ABI of libb200fem.so: the B200-native (sm_100a) partial-assembly
 * operator path behind MFEM's operator API.
 *
 * MFEM has no C ABI of its own: its extension seams are C++ virtual interfaces
 * (SURVEY.md section 8b).  Each entry point below names the reference interface
 * (file:line under the MFEM source tree, v4.9.1-dev) it replaces; the MFEM-side adapter
 * classes that bind them (`B200PAOperator : mfem::Operator`, ...) are in adapter/ and
 * INTEGRATION.md.
 *
 * Conventions
 *  - every function returns B2F_OK (0) or a B2F_ERR_* code; b2f_last_error() gives the text.
 *    No exceptions cross the boundary, nothing aborts the process.
 *  - pointers are DEVICE pointers unless the parameter name ends in `_h` (host memory).
 *  - tensors use the reference's layout: first index fastest, int32 indices, double values
 *    (linalg/dtensor.hpp:32; config/config.hpp real_t = double).
 *  - all work is enqueued on the context's stream (a cudaStream_t passed as void*;
 *    NULL = the legacy default stream the reference uses, general/forall.hpp:761) and is
 *    asynchronous unless stated otherwise.
 *  - the caller owns all vectors; handles own their device tables.
 *  - there is no CPU fallback: without a CUDA device b2f_ctx_create fails.
 */
#ifndef B200FEM_H
#define B200FEM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2F_VERSION 100

enum
{
   B2F_OK = 0,
   B2F_ERR_ARG = 1,         /* invalid argument                                   */
   B2F_ERR_CUDA = 2,        /* a CUDA runtime call or kernel launch failed        */
   B2F_ERR_UNSUPPORTED = 3, /* (dim, D1D, Q1D) or option without a kernel         */
   B2F_ERR_NCCL = 4,
   B2F_ERR_NO_DEVICE = 5    /* no usable CUDA device: the library has no CPU path */
};

typedef struct b2f_ctx b2f_ctx;                 /* device + stream + scratch              */
typedef struct b2f_restriction b2f_restriction; /* mfem::ElementRestriction               */
typedef struct b2f_pa b2f_pa;                   /* one BilinearFormIntegrator's PA data   */
typedef struct b2f_operator b2f_operator;       /* PABilinearFormExtension (+Constrained) */
typedef struct b2f_halo b2f_halo;               /* (Device)ConformingProlongationOperator */
typedef struct b2f_lor b2f_lor;                 /* BatchedLORAssembly's assembled SparseMatrix */

int b2f_version(void);
const char *b2f_last_error(void);

/* ---- context: mfem::Device("cuda") (general/device.hpp:128, device.cpp:191,554) ---- */
int b2f_ctx_create(int device, void *stream, b2f_ctx **out);
void b2f_ctx_destroy(b2f_ctx *ctx);
int b2f_ctx_sync(b2f_ctx *ctx);                 /* MFEM_STREAM_SYNC (general/cuda.hpp)    */
int b2f_ctx_device(const b2f_ctx *ctx);
int b2f_ctx_sm_count(const b2f_ctx *ctx);
/* Launch counter: number of kernels of THIS library launched since creation / reset. */
int64_t b2f_ctx_launch_count(const b2f_ctx *ctx);
void b2f_ctx_reset_launch_count(b2f_ctx *ctx);

/* CUDA-event timing of the PA operator kernel (pa_apply) launches alone, on the context's
 * stream: enable, run, then read the summed device time and launch count (synchronises). */
int b2f_ctx_kernel_timing(b2f_ctx *ctx, int enable);
int b2f_ctx_kernel_time(b2f_ctx *ctx, double *ms_total_h, int *launches_h);

/* ---- raw device memory for hosts without the CUDA runtime (mem_manager.hpp roles) --- */
int b2f_malloc(b2f_ctx *ctx, size_t bytes, void **dptr);
int b2f_free(b2f_ctx *ctx, void *dptr);
int b2f_malloc_host(b2f_ctx *ctx, size_t bytes, void **hptr); /* pinned */
int b2f_free_host(b2f_ctx *ctx, void *hptr);
int b2f_memcpy_h2d(b2f_ctx *ctx, void *dst, const void *src_h, size_t bytes); /* async */
int b2f_memcpy_d2h(b2f_ctx *ctx, void *dst_h, const void *src, size_t bytes); /* async */
int b2f_memcpy_d2d(b2f_ctx *ctx, void *dst, const void *src, size_t bytes);
int b2f_memset(b2f_ctx *ctx, void *dst, int byte, size_t bytes);

/* ---- host-side discretisation tables (C++ host code, no device work) ---------------- */
/* QuadratureFunctions1D::GaussLegendre / GaussLobatto on [0,1] (fem/intrules.cpp:620,708) */
int b2f_gauss_legendre_h(int np, double *pts_h, double *wts_h);
int b2f_gauss_lobatto_h(int np, double *pts_h, double *wts_h);
/* DofToQuad, TENSOR mode, of the order-p H1 Gauss-Lobatto basis at nq points
 * (fem/fe/fe_base.cpp:2697-2740, Poly_1D::Basis::Eval :1947): B_h,G_h are (nq, p+1). */
int b2f_h1_dof_to_quad_h(int p, int nq, const double *qpts_h, double *B_h, double *G_h);
/* IntegrationRule tensor weights (fem/intrules.cpp:43-96): W_h has nq^dim entries. */
int b2f_tensor_weights_h(int dim, int nq, const double *w1_h, double *W_h);
/* Vector::Randomize(seed) (linalg/vector.cpp:955-966): srand(seed), v[i] = rand()/(RAND_MAX+1.0);
 * seed 0 draws the seed from time(), as the reference does.  Uses the C library's global generator. */
int b2f_vector_randomize_h(double *v_h, int64_t n, int seed);

/* Structured hex/quad mesh + H1 space with the reference's numbering
 * (Mesh::MakeCartesian2D/3D mesh/mesh.cpp:3841,4388 incl. the space-filling-curve element
 * order; edge/face numbering mesh.cpp:8526-8560,8996; H1 dofs fem/fespace.cpp:3428-3536;
 * lexicographic dof_map fem/fe/fe_base.cpp:2497-2680).  Two-call pattern: pass NULL arrays
 * to get sizes.  elem_dof_h is the element-to-dof table in the element's NATIVE order
 * (what FiniteElementSpace::GetElementToDofTable().GetJ() holds); dof_map_h the
 * lexicographic -> native map; bdr_dofs_h the boundary (essential) dofs, sorted. */
int b2f_cartesian_h1_space_h(int dim, int nx, int ny, int nz, int p, int sfc_ordering,
                             int *ne, int *ndofs, int *elem_dof_h, int *dof_map_h,
                             int *nbdr, int *bdr_dofs_h,
                             double *elem_vertices_h /* (2^dim, dim, NE) lexicographic, or NULL */);

/* Cartesian block partition of that mesh over a px x py x pz process grid and the rank's
 * parallel H1 space tables -- what ParMesh + ParFiniteElementSpace + GroupCommunicator
 * produce in the reference (Mesh::CartesianPartitioning mesh/mesh.cpp:9188; group master =
 * lowest rank general/communication.cpp:120-124; true dofs fem/pfespace.cpp:2169; neighbour
 * tables general/communication.cpp:543,579).  No communication: every rank calls it with
 * its own `rank` (= ix + px*(iy + py*iz)).  All array arguments may be NULL;
 * sizes_h[11] = {lnx, lny, lnz, ne, n_ldofs, n_tdofs, n_nbr, n_shr, n_ext, n_bdr, n_interior};
 * the elements are ordered interior-first: [0, n_interior) touch no neighbour-owned dof.
 * The local space is numbered as b2f_cartesian_h1_space_h numbers the local block, then
 * permuted so that the rank's own (true) dofs come first: ltdof_ldof is the identity and a
 * true-dof vector is the head of a local-dof vector (the halo then works in place);
 * elem_vertices_h holds GLOBAL coordinates; bdr_ldofs_h the local dofs on the GLOBAL
 * boundary; ldof_lattice_h the point gx + LX*(gy + LY*gz) of each local dof on the global
 * (p*nx+1, p*ny+1, p*nz+1) lattice.  The remaining tables are b2f_halo_create's inputs. */
int b2f_cartesian_partition_h(int dim, int nx, int ny, int nz, int p, int sfc_ordering,
                              int px, int py, int pz, int rank, int *sizes_h,
                              int *elem_dof_h, int *dof_map_h, int *bdr_ldofs_h,
                              double *elem_vertices_h, int *ltdof_ldof_h, int *nbr_rank_h,
                              int *shr_offsets_h, int *shr_ltdof_h, int *ext_offsets_h,
                              int *ext_ldof_h, int64_t *ldof_lattice_h);

/* ---- ElementRestriction (fem/restriction.hpp:40, restriction.cpp:26-228) ------------ */
/* From the reference's own tables (ElementRestriction::GatherMap/Offsets/Indices,
 * restriction.hpp:104-106).  Tables are copied to the device. */
int b2f_restriction_create(b2f_ctx *ctx, int ne, int dof, int ndofs, int vdim, int byvdim,
                           const int *gather_map_h, const int *offsets_h,
                           const int *indices_h, b2f_restriction **out);
/* ElementRestriction::ElementRestriction (restriction.cpp:26-114): builds gather_map,
 * offsets and indices on the device from the element-to-dof table (signed, native order)
 * and the lexicographic dof_map (NULL = native ordering). */
int b2f_restriction_create_from_elem_dof(b2f_ctx *ctx, int ne, int dof, int ndofs, int vdim,
                                         int byvdim, const int *elem_dof_h,
                                         const int *dof_map_h, b2f_restriction **out);
void b2f_restriction_destroy(b2f_restriction *r);
/* device pointers to the tables (for parity checks and for adapters) */
int b2f_restriction_tables(const b2f_restriction *r, const int **gather_map,
                           const int **offsets, const int **indices);
/* ElementRestriction::Mult (restriction.cpp:116): L -> E, bit-exact */
int b2f_restriction_mult(b2f_restriction *r, const double *x_l, double *y_e);
/* ElementRestriction::AbsMult (restriction.cpp:138) */
int b2f_restriction_abs_mult(b2f_restriction *r, const double *x_l, double *y_e);
/* ElementRestriction::MultTranspose / AddMultTranspose (restriction.cpp:159-201):
 * E -> L, per-dof sequential sum in ascending E-index order, bit-exact. */
int b2f_restriction_mult_transpose(b2f_restriction *r, const double *x_e, double *y_l, int add);
/* ElementRestriction::AbsMultTranspose (restriction.cpp:203) */
int b2f_restriction_abs_mult_transpose(b2f_restriction *r, const double *x_e, double *y_l);

/* ---- BilinearFormIntegrator PA seam (fem/bilininteg.hpp:49-95) ---------------------- */
/* DiffusionIntegrator::AssemblePA (fem/integ/bilininteg_diffusion_pa.cpp:108) ->
 * PADiffusionSetup (bilininteg_diffusion_kernels.cpp:107-386).  J is the reference's
 * GeometricFactors::J, (Q^dim, dim, dim, NE).  coeff: coeff_dim = 1 (scalar), dim (vector),
 * dim(dim+1)/2 (symmetric matrix) or dim*dim; coeff_size == coeff_dim means constant,
 * otherwise (coeff_dim, Q^dim, NE).  B_h,G_h: (Q1D, D1D); W_h: Q1D^dim weights. */
int b2f_diffusion_setup(b2f_ctx *ctx, int dim, int d1d, int q1d, int64_t ne,
                        const double *B_h, const double *G_h, const double *W_h,
                        const double *J, const double *coeff, int coeff_dim,
                        int64_t coeff_size, b2f_pa **out);
/* MassIntegrator::AssemblePA (fem/integ/bilininteg_mass_pa.cpp:25-91): D = W c detJ.
 * coeff_size == 1 means constant, otherwise (Q^dim, NE). */
int b2f_mass_setup(b2f_ctx *ctx, int dim, int d1d, int q1d, int64_t ne, const double *B_h,
                   const double *W_h, const double *detJ, const double *coeff,
                   int64_t coeff_size, b2f_pa **out);
/* Same two integrators with the geometric factors (mesh/mesh.cpp:15420-15498,
 * fem/qinterp/grad.hpp:235, det.hpp) fused in: J is evaluated from the nodal coordinates
 * E-vector (node_d1d^dim, dim, NE) on the fly and never written to memory.
 * Bn_h,Gn_h: (Q1D, node_d1d) tables of the nodal basis at the quadrature points.
 * kind: 0 = diffusion, 1 = mass.  Scalar constant coefficient. */
int b2f_pa_setup_from_nodes(b2f_ctx *ctx, int kind, int dim, int d1d, int q1d, int64_t ne,
                            const double *B_h, const double *G_h, const double *W_h,
                            int node_d1d, const double *Bn_h, const double *Gn_h,
                            const double *nodes_e, double coeff, b2f_pa **out);
void b2f_pa_destroy(b2f_pa *pa);
/* the quadrature-point data ("pa_data", bilininteg.hpp:2287): device pointer and length */
int b2f_pa_data(const b2f_pa *pa, const double **data, int64_t *size);
/* BilinearFormIntegrator::AddMultPA (bilininteg.hpp:78): y_e += A_e x_e on E-vectors.
 * Replaces (Smem)PADiffusionApply2D/3D (bilininteg_diffusion_kernels.hpp:514-1215) and
 * (Smem)PAMassApply2D/3D (bilininteg_mass_kernels.hpp:440-1165). */
int b2f_pa_add_mult(b2f_pa *pa, const double *x_e, double *y_e);
/* BilinearFormIntegrator::AddMultTransposePA (bilininteg.hpp:88): the same kernel for the
 * symmetric operators; like the reference (bilininteg_diffusion_pa.cpp:95-106) it fails for a
 * diffusion integrator with a non-symmetric matrix coefficient. */
int b2f_pa_add_mult_transpose(b2f_pa *pa, const double *x_e, double *y_e);
/* BilinearFormIntegrator::AddAbsMultPA (DiffusionIntegrator: bilininteg_diffusion_pa.cpp:196-209,
 * MassIntegrator: bilininteg_mass_pa.cpp:207-228): the same kernels with |B|, |G| and
 * |pa_data| (an absolute-value twin of the integrator is built on first use). */
int b2f_pa_add_abs_mult(b2f_pa *pa, const double *x_e, double *y_e);
/* VectorDiffusionIntegrator / VectorMassIntegrator::AddMultPA (fem/integ/
 * bilininteg_vecdiffusion_pa.cpp, bilininteg_vecmass_pa.cpp; CEED BP2/4/6): the same operator on
 * every component of an E-vector laid out (dofs, vdim, NE); one quadrature-data load serves
 * all components. */
int b2f_pa_add_mult_vec(b2f_pa *pa, int vdim, const double *x_e, double *y_e);
/* BilinearFormIntegrator::AssembleDiagonalPA (bilininteg.hpp:68): diag_e += diag(A_e).
 * Replaces (Smem)PADiffusionDiagonal2D/3D, (Smem)PAMassAssembleDiagonal2D/3D. */
int b2f_pa_assemble_diagonal(b2f_pa *pa, double *diag_e);

/* ---- DomainLFIntegrator fast assembly (LinearForm::UseFastAssembly) ------------------- */
/* DomainLFIntegrator / VectorDomainLFIntegrator::AssembleDevice (fem/integ/
 * lininteg_domain.cpp:21-72) -> DLFEvalAssemble2D/3D (lininteg_domain_kernels.hpp:76-297):
 * y_e(d, c, e) += sum_q B(q,d) W(q) f_c(q,e) detJ(q,e) on E-vectors (D1D^dim, vdim, NE).
 * markers (device, [NE], 0 = skip the element) may be NULL; detJ (device, (Q^dim, NE)) as
 * GeometricFactors::detJ; coeff (device) is the CoefficientVector: vdim constants
 * (coeff_size == vdim) or (vdim, Q^dim, NE).  LinearFormExtension::Assemble
 * (fem/linearform_ext.cpp:20-66) then applies b2f_restriction_mult_transpose. */
int b2f_domain_lf_assemble(b2f_ctx *ctx, int dim, int d1d, int q1d, int vdim, int64_t ne,
                           const int *markers, const double *B_h, const double *W_h,
                           const double *detJ, const double *coeff, int64_t coeff_size,
                           double *y_e);

/* ---- PABilinearFormExtension + ConstrainedOperator ---------------------------------- */
/* PABilinearFormExtension (fem/bilinearform_ext.cpp:256-564) over one restriction and
 * n integrators.  A restriction with vdim > 1 (byNODES or byVDIM) makes it the vector operator
 * of Vector{Mass,Diffusion}Integrator on vdof vectors (rank-parallel with byVDIM, see
 * b2f_halo_create); ess_ldofs_h (may be NULL / 0) is the essential dof list that
 * FormLinearSystem wraps around it as a ConstrainedOperator (linalg/operator.cpp:526). */
int b2f_operator_create(b2f_ctx *ctx, b2f_restriction *r, b2f_pa *const *integrators,
                        int n_integrators, const int *ess_dofs_h, int n_ess,
                        b2f_operator **out);
void b2f_operator_destroy(b2f_operator *op);
/* PABilinearFormExtension::Mult (bilinearform_ext.cpp:487-564): y = R^T (sum_i A_i) R x as
 * ONE fused kernel per integrator (gather, contraction, D, transpose, scatter): no
 * E-vector in memory.  1e-12 relative parity (atomic scatter order is not fixed). */
int b2f_operator_mult(b2f_operator *op, const double *x_l, double *y_l);
/* The reference's unfused sequence (R, AddMultPA, R^T) with E-vectors in memory; the
 * scatter is deterministic.  For parity cross-checks. */
int b2f_operator_mult_unfused(b2f_operator *op, const double *x_l, double *y_l);
/* ConstrainedOperator::Mult (linalg/operator.cpp:601-660, DIAG_ONE):
 * z = x; z[ess] = 0; y = A z; y[ess] = x[ess]. */
int b2f_operator_mult_constrained(b2f_operator *op, const double *x_l, double *y_l);
/* Operator::AbsMult (linalg/operator.hpp:96): PABilinearFormExtension::AbsMult =
 * MultInternal(x, y, useAbs = true) (bilinearform_ext.cpp:487-564: |R|, AddAbsMultPA with
 * markers, |R|^T), wrapped by RAPOperator::AbsMult when a halo is attached, and
 * ConstrainedOperator::AbsMult (linalg/operator.cpp:663-735).  Scalar spaces. */
int b2f_operator_abs_mult(b2f_operator *op, const double *x, double *y);
int b2f_operator_abs_mult_constrained(b2f_operator *op, const double *x, double *y);
/* Operator::DiagonalPolicy of the ConstrainedOperator (linalg/operator.hpp:47-53,
 * operator.cpp:526-660): 0 = DIAG_ZERO, 1 = DIAG_ONE (default, what FormLinearSystem uses),
 * 2 = DIAG_KEEP (as in the reference, Mult then fails when there are constraints). */
int b2f_operator_set_diag_policy(b2f_operator *op, int policy);
/* ConstrainedOperator::AssembleDiagonal (operator.cpp:543-572) */
int b2f_operator_assemble_diagonal_constrained(b2f_operator *op, double *diag);
/* PABilinearFormExtension::AssembleDiagonal (bilinearform_ext.cpp:370-454) */
int b2f_operator_assemble_diagonal(b2f_operator *op, double *diag_l);
/* ConstrainedOperator::EliminateRHS (linalg/operator.cpp:574-599) */
int b2f_operator_eliminate_rhs(b2f_operator *op, const double *x_l, double *b_l);
/* Attach the rank-parallel prolongation: the operator becomes
 * RAPOperator(P, A, P) (linalg/operator.hpp:981-1033) on true-dof vectors. */
int b2f_operator_set_halo(b2f_operator *op, b2f_halo *halo);
int b2f_operator_height(const b2f_operator *op); /* true-dof size if a halo is set */
/* BilinearForm::AddDomainIntegrator(bfi, elem_marker) (fem/bilinearform.hpp; applied by
 * PABilinearFormExtension::AddMultWithMarkers / AssembleDiagonal, bilinearform_ext.cpp:
 * 374-399, 753-847): integrator `integrator` acts only on elements e with
 * markers_h[elem_attributes_h[e] - 1] != 0 (attributes are 1-based as in Mesh::GetAttribute).
 * markers_h = NULL removes the restriction.  The enabled elements are compacted at this call,
 * so the fused operator streams only them; the diagonal follows the reference literally
 * (the E-diagonal accumulated so far is reset on the integrator's unmarked elements). */
int b2f_operator_set_markers(b2f_operator *op, int integrator, const int *elem_attributes_h,
                             const int *markers_h, int n_markers);
/* Declare that elements [0, n_interior) touch no dof owned by another rank (the partition
 * generator orders elements that way).  With a peer-memory halo the operator then overlaps
 * the shared-dof exchange with interior-element work (the reference overlaps its local copy
 * with MPI_Isend/Irecv, pfespace.cpp:5786-5813).  -1 = unknown (no overlap). */
int b2f_operator_set_interior(b2f_operator *op, int64_t n_interior);

/* ---- OperatorJacobiSmoother (linalg/solvers.cpp:324-450) ----------------------------- */
/* Setup: dinv = damping / diag; dinv[ess] = damping (solvers.cpp:401-425). */
int b2f_jacobi_setup(b2f_ctx *ctx, int n, const double *diag, const int *ess_dofs /*device*/,
                     int n_ess, double damping, double *dinv);
/* Mult (iterative_mode = false): y = dinv .* x */
int b2f_jacobi_mult(b2f_ctx *ctx, int n, const double *dinv, const double *x, double *y);

/* ---- Vector BLAS-1 (linalg/vector.cpp:436,570,1079; general/reducers.hpp:536) -------- */
int b2f_vector_dot(b2f_ctx *ctx, int64_t n, const double *x, const double *y, double *result_h);
/* z = x + a*y */
int b2f_vector_add(b2f_ctx *ctx, int64_t n, const double *x, double a, const double *y, double *z);

/* ---- CGSolver (linalg/solvers.cpp:869-1050) ----------------------------------------- */
/* PCG with the reference's exact control flow (iterative_mode = false: x0 = 0):
 * r0 = max(nom0 rel_tol^2, abs_tol^2); stop when (B r, r) <= r0 or after max_iter.
 * dinv == NULL: no preconditioner.  The operator is applied as ConstrainedOperator (and
 * RAPOperator when a halo is attached); dot products are NCCL all-reduced across ranks
 * when a halo is attached (IterativeSolver::Dot, solvers.cpp:58).
 * All scalars live on the device; the host is consulted once per `check_every`
 * iterations (1 = the reference's behaviour of testing convergence every iteration).
 * history_h (may be NULL) receives (B r, r) for iterations 0..iters (capacity entries). */
typedef struct
{
   double rel_tol, abs_tol;
   int max_iter;
   int check_every;
} b2f_pcg_options;
typedef struct
{
   int iterations;
   int converged;
   double initial_norm, final_norm; /* sqrt((B r, r)) */
} b2f_pcg_result;
int b2f_pcg(b2f_operator *op, const double *dinv, const double *b, double *x,
            const b2f_pcg_options *opt, b2f_pcg_result *result_h, double *history_h,
            int history_capacity);

/* ---- OperatorChebyshevSmoother (linalg/solvers.hpp:392-505, solvers.cpp:453-658) ------
 * Polynomial smoother z = p(D^-1 A) D^-1 r of order 1..5 on the eigenvalue window
 * [0.3, 1.2] max_eig_estimate; A is the constrained operator of `op`, D its diagonal with
 * the essential entries replaced by 1 (Setup(), solvers.cpp:538-550). */
typedef struct b2f_chebyshev b2f_chebyshev;
/* The order coefficients of Setup() (solvers.cpp:553-618); host only, no device needed.
 * Fails for order outside 1..5 like the reference's MFEM_ABORT. */
int b2f_chebyshev_coefficients_h(int order, double max_eig_estimate, double *coeffs_h);
/* PowerMethod::EstimateLargestEigenvalue (linalg/operator.cpp:886-945) on
 * ProductOperator(OperatorJacobiSmoother(diag, ess, 1.0), A) as the smoother's second
 * constructor runs it (solvers.cpp:496-513).  dinv = b2f_jacobi_setup(..., damping 1);
 * v0 (device, true-dof size) holds the start vector on entry — the caller draws it
 * (Vector::Randomize(seed), linalg/vector.cpp:955) — and the last iterate on return.
 * Dot products are all-reduced when the operator has a halo. */
int b2f_power_method(b2f_operator *op, const double *dinv, double *v0, int num_steps,
                     double tolerance, double *eigenvalue_h);
int b2f_chebyshev_create(b2f_operator *op, const double *diag, int order,
                         double max_eig_estimate, b2f_chebyshev **out);
void b2f_chebyshev_destroy(b2f_chebyshev *cheb);
/* Mult (solvers.cpp:620-658; iterative_mode is not supported there either): order-1
 * operator applications, each diagonal scaling fused with the accumulation into y. */
int b2f_chebyshev_mult(b2f_chebyshev *cheb, const double *x, double *y);
/* CGSolver with the Chebyshev smoother as preconditioner (same loop as b2f_pcg). */
int b2f_pcg_chebyshev(b2f_operator *op, b2f_chebyshev *cheb, const double *b, double *x,
                      const b2f_pcg_options *opt, b2f_pcg_result *result_h,
                      double *history_h, int history_capacity);

/* ---- rank-parallel prolongation: ConformingProlongationOperator ----------------------
 * (fem/pfespace.hpp:615-697, pfespace.cpp:5479-5927) with the shared-dof exchange as
 * NCCL send/recv over NVLink instead of MPI_Isend/Irecv. */
/* rank 0 creates the id and distributes it out of band (e.g. torch.distributed) */
int b2f_nccl_unique_id_h(void *id128_h);
/* For a vector space (vdim > 1, Ordering::byVDIM) every table is on vdofs (vdof = ldof*vdim
 * + c), as the communicator of a ParFiniteElementSpace with vdim > 1 holds them; n_ldofs and
 * n_tdofs then count vdofs. */
int b2f_halo_create(b2f_ctx *ctx, const void *nccl_id128_h, int rank, int nranks,
                    int n_ldofs, int n_tdofs,
                    const int *ltdof_ldof_h,         /* [n_tdofs] owned ldofs, ascending   */
                    int n_nbr, const int *nbr_rank_h,
                    const int *shr_offsets_h,        /* [n_nbr+1] into shr_ltdof           */
                    const int *shr_ltdof_h,          /* owned tdofs sent to each neighbour */
                    const int *ext_offsets_h,        /* [n_nbr+1] into ext_ldof            */
                    const int *ext_ldof_h,           /* ldofs owned by each neighbour      */
                    b2f_halo **out);
void b2f_halo_destroy(b2f_halo *h);
/* The ParFiniteElementSpace binding: b2f_halo_create's tables from the reference's own group
 * tables -- GroupTopology (master rank and the group's number on the master, member ranks),
 * GroupCommunicator::GroupLDofTable and the ldof -> ltdof map (-1 = not owned).  Restates
 * GroupCommunicator::Finalize / GetNeighborLTDofTable / GetNeighborLDofTable
 * (general/communication.cpp:368-612) and the table part of DeviceConformingProlongationOperator's
 * constructor (fem/pfespace.cpp:5625-5690); host code, no MPI (include/b200fem_groups.hpp is the
 * header-only C++ form an MPI adapter calls).  Group 0 (the local group) is ignored.  Two-call
 * pattern: array arguments may be NULL; sizes_h[4] = {n_tdofs, n_nbr, n_shr, n_ext}. */
int b2f_halo_tables_from_groups_h(int my_rank, int n_groups, const int *group_master_rank_h,
                                  const int *group_master_group_h, const int *group_rank_offsets_h,
                                  const int *group_ranks_h, const int *group_ldof_offsets_h,
                                  const int *group_ldofs_h, int n_ldofs, const int *ldof_ltdof_h,
                                  int *sizes_h, int *ltdof_ldof_h, int *nbr_rank_h, int *shr_offsets_h,
                                  int *shr_ltdof_h, int *ext_offsets_h, int *ext_ldof_h);
/* Peer-memory exchange (one process per GPU on one NVLink/NVSwitch node): instead of
 * pack + MPI_Isend/Irecv (pfespace.cpp:5786-5803) or pack + ncclSend/ncclRecv, the pack
 * kernel stores straight into the neighbour's receive buffer through a CUDA-IPC mapping and
 * publishes an epoch flag that the neighbour's unpack / assembly kernel waits on: the
 * exchange is two small kernels on the solver's stream, no library call, no host sync.
 * Every rank exports a blob of b2f_halo_p2p_blob_bytes() bytes, the caller gathers the blobs
 * of all ranks (MPI_Allgather / torch.distributed) into blobs_h[rank] and connects.  Used by
 * the in-place P / P^T of the operator (owned-first numbering). */
int b2f_halo_p2p_blob_bytes(void);
int b2f_halo_p2p_export(b2f_halo *h, void *blob_h);
int b2f_halo_p2p_connect(b2f_halo *h, const void *blobs_h, int nranks);
/* P: true-dof vector -> local-dof vector (pfespace.cpp:5760-5816) */
int b2f_halo_mult(b2f_halo *h, const double *x_t, double *y_l);
/* P^T: local-dof vector -> true-dof vector, fixed summation order (pfespace.cpp:5818-5927) */
int b2f_halo_mult_transpose(b2f_halo *h, const double *x_l, double *y_t);
/* InnerProduct(comm, x, y) (linalg/vector.hpp:773): local dot + all-reduce */
int b2f_halo_dot(b2f_halo *h, int64_t n, const double *x, const double *y, double *result_h);


/* ---- LOR (low-order-refined) batched assembly, H1, dim = 2 or 3 --------------------------- */
/* BatchedLORAssembly::FormLORVertexCoordinates (fem/lor/lor_batched.cpp:75-112): the mesh nodes
 * (device E-vector (nd^dim, dim, NE), nd = node_nd1d) evaluated at the (p+1)^dim Gauss-Lobatto points
 * of every element: x_vert (device) is (dim, (p+1)^dim, NE), byVDIM.  Bn_h (host) is the 1D nodal basis of
 * the mesh at those points, (p+1, nd) row-major. */
int b2f_lor_vertex_coordinates(b2f_ctx *ctx, int dim, int p, int node_nd1d, int64_t ne, const double *Bn_h,
                               const double *nodes_e, double *x_vert);
/* BatchedLORAssembly::Assemble for an H1 space with a MassIntegrator and / or DiffusionIntegrator
 * (fem/lor/lor_batched.cpp:712-734): BatchedLOR_H1::Assemble2D / Assemble3D (fem/lor/lor_h1_impl.hpp:
 * 19-182 / 184-455),
 * FillI / FillJAndData (lor_batched.cpp:143-369), SparseMatrix::EliminateBC(ess, DIAG_KEEP)
 * (linalg/sparsemat.cpp:2375).  r is the space's lexicographic ElementRestriction; c_mass / c_diff
 * (device) are the projected coefficients c1 / c2 of BatchedLORKernel (lor_batched.hpp:180-216): one
 * value (n == 1; 0 when the integrator is absent) or ((p+1)^dim, NE); ess (device, may be NULL when
 * n_ess == 0) lists the essential dofs.  Rows and the columns inside a row come in the order of the
 * reference's serial host loop.  keep_sparse_ij != 0 keeps the macro-element stencil array
 * (BatchedLORAssembly::sparse_ij, (3^dim, (p+1)^dim, NE)) for b2f_lor_sparse_ij. */
int b2f_lor_assemble_h1(b2f_ctx *ctx, const b2f_restriction *r, int dim, int p, const double *x_vert,
                        const double *c_mass, int64_t n_mass, const double *c_diff, int64_t n_diff,
                        const int *ess, int n_ess, int keep_sparse_ij, b2f_lor **out);
/* the assembled SparseMatrix: device pointers I (n_rows + 1), J (nnz), A (nnz); any may be NULL */
int b2f_lor_csr(const b2f_lor *lor, int *n_rows, int64_t *nnz, const int **I, const int **J, const double **A);
int b2f_lor_sparse_ij(const b2f_lor *lor, const double **V, int64_t *size);
/* device time of the kernels of the assembly that produced `lor` (CUDA events on the context's stream
 * around stencil array + count + scan and around fill + elimination; allocations excluded) */
int b2f_lor_kernel_ms(const b2f_lor *lor, double *ms);
/* SparseMatrix::Mult: y = A x (device vectors) */
int b2f_lor_mult(b2f_lor *lor, const double *x, double *y);
void b2f_lor_destroy(b2f_lor *lor);

#ifdef __cplusplus
}
#endif
#endif /* B200FEM_H */
