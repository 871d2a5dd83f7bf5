/* oracle/oracle_pa.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's (MFEM v4.9.1-dev) partial-assembly hot
 * path.  Each function cites the reference file:line it follows.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (mfem_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below
 * against fixtures dumped from the unmodified reference by oracle/ref_driver.cpp
 * (tests/golden/<case>.npz, generator: oracle/make_golden.py) and against the golden
 * norms recorded in BASELINE.md section 2.
 *
 * Conventions (the reference's): column-major "first index fastest" tensors
 * (linalg/dtensor.hpp:32), int32 indices, double precision.
 *   B,G     : (Q1D, D1D)            B[q + Q1D*d]
 *   E-vector: (D1D^dim, NE)         lexicographic, element-major
 *   J       : (Q^dim, sdim, dim, NE)   J(q,i,j,e) = d x_i / d xi_j
 *   D diff  : (Q^dim, 3|6 (sym) or 4|9, NE);  D mass: (Q^dim, NE)
 */
#ifndef ORACLE_PA_H
#define ORACLE_PA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* fem/restriction.cpp:26-114 */
void orc_restriction_build(int ne, int dof, int ndofs, const int *elem_dof,
                           const int *dof_map /* may be NULL */, int *gather_map,
                           int *offsets /* ndofs+1 */, int *indices);
/* fem/restriction.cpp:116-136 (vdim = 1) */
void orc_restriction_mult(int ne, int dof, const int *gather_map, const double *x,
                          double *y);
/* fem/restriction.cpp:159-201 ; add != 0 => AddMultTranspose */
void orc_restriction_mult_transpose(int ndofs, int dof, const int *offsets,
                                    const int *indices, const double *x, double *y,
                                    int add);
/* fem/restriction.cpp:203-228 */
void orc_restriction_abs_mult_transpose(int ndofs, int dof, const int *offsets,
                                        const int *indices, const double *x, double *y);

/* fem/intrules.cpp:620-690 (Gauss-Legendre on [0,1]) and :708-830 (Gauss-Lobatto) */
void orc_gauss_legendre(int np, double *pts, double *wts);
void orc_gauss_lobatto(int np, double *pts, double *wts);
/* fem/fe/fe_base.cpp:1811-1850 (barycentric weights), :1947-2000 (Eval u,d) and
 * :2697-2740 (GetTensorDofToQuad): B,G are (nq, nd) column-major. */
void orc_basis_tables(int nd, const double *nodes, int nq, const double *qpts,
                      double *B, double *G);
/* fem/intrules.cpp:43-96: tensor weights W[qx + Q*(qy + Q*qz)] = wx*wy*wz */
void orc_tensor_weights(int dim, int nq, const double *w1, double *W);

/* mesh/mesh.cpp:15420-15498 + fem/qinterp/grad.hpp:235 (Derivatives, byNODES E-vector
 * (nd^dim, dim, NE)) + fem/qinterp/det.hpp: J and detJ at the quadrature points. */
void orc_geometric_factors(int dim, int nd1d, int q1d, int64_t ne, const double *Bn,
                           const double *Gn, const double *nodesE, double *J,
                           double *detJ /* may be NULL */);

/* fem/integ/bilininteg_diffusion_kernels.cpp:146-215 (2D), :262-386 (3D).
 * coeff_dim: 1 scalar, dim vector, dim(dim+1)/2 sym. matrix, dim*dim full matrix.
 * coeff_const != 0 => C has coeff_dim entries, else (coeff_dim, Q^dim, NE). */
void orc_diffusion_setup(int dim, int q1d, int coeff_dim, int coeff_const, int64_t ne,
                         const double *W, const double *J, const double *C, double *D);
/* fem/integ/bilininteg_mass_pa.cpp:73-90 */
void orc_mass_setup(int nq, int64_t ne, const double *W, const double *detJ,
                    const double *C, int coeff_const, double *D);

/* fem/integ/bilininteg_diffusion_kernels.hpp:514-633 (2D), :796-986 (3D): Y += G^T D G X */
void orc_diffusion_apply(int dim, int d1d, int q1d, int64_t ne, int symmetric,
                         const double *B, const double *G, const double *D,
                         const double *X, double *Y);
/* fem/integ/bilininteg_mass_kernels.hpp:440-536 (2D), :664-813 (3D): Y += B^T D B X */
void orc_mass_apply(int dim, int d1d, int q1d, int64_t ne, const double *B,
                    const double *D, const double *X, double *Y);
/* fem/integ/bilininteg_diffusion_kernels.hpp:92-155 (2D), :268-366 (3D) */
void orc_diffusion_diag(int dim, int d1d, int q1d, int64_t ne, int symmetric,
                        const double *B, const double *G, const double *D, double *Y);
/* fem/integ/bilininteg_mass_kernels.hpp:28-80 (2D), :228-330 (3D) */
void orc_mass_diag(int dim, int d1d, int q1d, int64_t ne, const double *B,
                   const double *D, double *Y);

/* A PA operator on L-vectors: y = R^T A_E R x  (fem/bilinearform_ext.cpp:487-564),
 * optionally wrapped as ConstrainedOperator (linalg/operator.cpp:601-660). */
typedef struct
{
   int dim, d1d, q1d, is_mass, symmetric;
   int64_t ne;
   int ndofs;
   const int *gather_map, *offsets, *indices;
   const double *B, *G, *D;
   const int *ess; /* essential (constrained) dofs, may be NULL */
   int ness;
   double *xe, *ye, *z; /* work: ne*dof, ne*dof, ndofs */
} orc_operator;

void orc_op_mult(const orc_operator *op, const double *x, double *y);            /* unconstrained */
void orc_op_mult_constrained(const orc_operator *op, const double *x, double *y);
void orc_op_diag(const orc_operator *op, double *diag);                          /* bilinearform_ext.cpp:370-454 */
/* linalg/operator.cpp:574-599 EliminateRHS:  b -= A w (w = x on ess, 0 elsewhere); b[ess] = x[ess] */
void orc_op_eliminate_rhs(const orc_operator *op, const double *x, double *b);

/* linalg/solvers.cpp:401-425 */
void orc_jacobi_setup(int n, const double *diag, const int *ess, int ness, double damping,
                      double *dinv);
/* linalg/solvers.cpp:869-1050. dinv == NULL => unpreconditioned CG.  history (may be
 * NULL) receives (B r, r) for iterations 0..iters.  Returns number of iterations. */
int orc_pcg(const orc_operator *op, const double *dinv, const double *b, double *x,
            double rel_tol, double abs_tol, int max_iter, int *converged,
            double *final_norm, double *history, int history_cap);

/* linalg/vector.cpp:955-966: srand(seed); v[i] = rand()/(RAND_MAX+1.0) */
/* DomainLFIntegrator device assembly (fem/integ/lininteg_domain_kernels.hpp:76-297) */
void orc_domain_lf(int dim, int d, int q, int vdim, int ne, const int *markers, const double *B,
                   const double *W, const double *detJ, const double *coeff, int cst, double *y);
/* OperatorChebyshevSmoother (linalg/solvers.cpp:453-658), PowerMethod (linalg/operator.cpp:886) */
int orc_chebyshev_coeffs(int order, double max_eig, double *coeffs);
void orc_chebyshev_mult(const orc_operator *op, const double *dinv, int order, const double *coeffs,
                        const double *x, double *y);
double orc_power_method(const orc_operator *op, const double *dinv, double *v0, int steps, double tol);
int orc_pcg_chebyshev(const orc_operator *op, const double *dinv, int order, const double *coeffs,
                      const double *b, double *x, double rel_tol, double abs_tol, int max_iter,
                      int *converged, double *final_norm, double *history, int history_cap);
void orc_randomize(int n, int seed, double *v);

#ifdef __cplusplus
}
#endif
#endif
