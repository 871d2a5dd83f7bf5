// mfem_b200/csrc/capi.cu -- the extern "C" surface of libb200fem.so (include/b200fem.h).
//
// Thin glue: argument checks, handle lifetime, and the host-side orchestration that the
// reference does in PABilinearFormExtension (fem/bilinearform_ext.cpp:256-564),
// ConstrainedOperator (linalg/operator.cpp:526-660), OperatorJacobiSmoother
// (linalg/solvers.cpp:324-450) and CGSolver::Mult (linalg/solvers.cpp:869-1050).
// No arithmetic happens on the host; without a CUDA device every entry point fails.
#include <cstdarg>
#include <cstring>

#include <climits>
#include <new>
#include "common.cuh"

namespace b2f
{

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(g_err, sizeof(g_err), fmt, ap);
   va_end(ap);
}

// restriction.cu
int restriction_build_device(b2f_restriction *r, const int *elem_dof_d, const int *dof_map_d);
int restriction_gather(b2f_restriction *r, const double *x, double *y, bool abs);
int restriction_scatter(b2f_restriction *r, const double *x, double *y, int mode);
// vector_ops.cu
int vector_add(b2f_ctx *ctx, int64_t n, const double *x, double a, const double *y, double *z);
int vector_mul(b2f_ctx *ctx, int64_t n, const double *a, const double *x, double *y);
int jacobi_setup(b2f_ctx *ctx, int n, const double *diag, const int *ess, int n_ess,
                 double damping, double *dinv);
int pcg_state_bytes();
int pcg_after_nom0(b2f_ctx *ctx, void *state, double rel_tol, double abs_tol, int max_iter,
                   double *history, int history_cap);
int pcg_after_den0(b2f_ctx *ctx, void *state);
int pcg_update(b2f_ctx *ctx, void *state, int64_t n, double *x, const double *d, double *r,
               const double *q, const double *dinv, double *z);
int pcg_after_betanom(b2f_ctx *ctx, void *state, double *history, int history_cap);
int pcg_direction(b2f_ctx *ctx, void *state, int64_t n, const double *z, double *d);
int lf_domain(b2f_ctx *ctx, int dim, int d, int q, int vdim, int64_t ne, const int *markers,
              const double *B_d, const double *W_d, const double *detJ, const double *coeff, int cst,
              double *y_e);
int cheb_step(b2f_ctx *ctx, int64_t n, const double *src, const double *dinv, double c, double *res,
              double *y, bool first);
int normalize_by(b2f_ctx *ctx, int64_t n, const double *d_normsq, double *v);
int pcg_after_den(b2f_ctx *ctx, void *state);
double *pcg_dot_slot(void *state);
void pcg_read_state(const void *host_copy, int *iter, int *done, int *converged, int *final_iter,
                    double *nom0, double *betanom);
// halo.cu
int halo_mult(b2f_halo *h, const double *x_t, double *y_l);
int halo_mult_transpose(b2f_halo *h, const double *x_l, double *y_t);
int halo_allreduce_sum(b2f_halo *h, double *d_scalar, int count);
int halo_sizes(const b2f_halo *h, int *n_ldofs, int *n_tdofs);
const int *halo_ldof_to_tdof_h(const b2f_halo *h); // host array [n_ldofs]: tdof or -1
bool halo_is_prefix(const b2f_halo *h);
int halo_mult_inplace(b2f_halo *h, double *v);
int halo_mult_transpose_inplace(b2f_halo *h, double *v);
bool halo_is_p2p(const b2f_halo *h);
int halo_p_push(b2f_halo *h, const double *v);
int halo_p_wait(b2f_halo *h, double *v);
int halo_pt_push(b2f_halo *h, const double *v);
int halo_pt_wait(b2f_halo *h, double *v);

template <typename T>
static int upload(const T *h, size_t n, T **d)
{
   *d = nullptr;
   B2F_CUDA(cudaMalloc((void **)d, sizeof(T) * (n ? n : 1)));
   if (n) { B2F_CUDA(cudaMemcpy(*d, h, sizeof(T) * n, cudaMemcpyHostToDevice)); }
   return B2F_OK;
}

} // namespace b2f

using namespace b2f;

#define B2F_TRY(expr)                                                                  \
   do                                                                                  \
   {                                                                                   \
      int rc_ = (expr);                                                                \
      if (rc_ != B2F_OK) { return rc_; }                                               \
   } while (0)

extern "C" {

int b2f_version(void) { return B2F_VERSION; }
const char *b2f_last_error(void) { return g_err; }

// ---- context ----------------------------------------------------------------------------
int b2f_ctx_create(int device, void *stream, b2f_ctx **out)
{
   B2F_REQUIRE(out, "b2f_ctx_create: out is NULL");
   *out = nullptr;
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0)
   {
      cudaGetLastError();
      set_error("b2f_ctx_create: no CUDA device (%s); libb200fem has no CPU path",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
      return B2F_ERR_NO_DEVICE;
   }
   B2F_REQUIRE(device >= 0 && device < ndev, "b2f_ctx_create: device %d out of range [0,%d)",
               device, ndev);
   B2F_CUDA(cudaSetDevice(device));
   b2f_ctx *ctx = new (std::nothrow) b2f_ctx;
   B2F_REQUIRE(ctx, "b2f_ctx_create: out of host memory");
   ctx->device = device;
   ctx->stream = (cudaStream_t)stream;
   auto init = [&]() -> int
   {
      B2F_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
      B2F_CUDA(cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin,
                                      device));
      ctx->n_partials = ctx->sm_count * 8;
      B2F_CUDA(cudaMalloc((void **)&ctx->d_partials, sizeof(double) * ctx->n_partials));
      B2F_CUDA(cudaMalloc((void **)&ctx->d_counter, sizeof(unsigned int)));
      B2F_CUDA(cudaMemset(ctx->d_counter, 0, sizeof(unsigned int)));
      B2F_CUDA(cudaMalloc((void **)&ctx->d_scalars, sizeof(double) * 64));
      B2F_CUDA(cudaMemset(ctx->d_scalars, 0, sizeof(double) * 64));
      B2F_CUDA(cudaMallocHost((void **)&ctx->h_scalars, sizeof(double) * 64));
      return B2F_OK;
   };
   if (const int rc = init())
   {
      b2f_ctx_destroy(ctx); // a failed creation leaks nothing
      return rc;
   }
   *out = ctx;
   return B2F_OK;
}

void b2f_ctx_destroy(b2f_ctx *ctx)
{
   if (!ctx) { return; }
   cudaSetDevice(ctx->device);
   cudaStreamSynchronize(ctx->stream);
   cudaFree(ctx->d_partials);
   cudaFree(ctx->d_counter);
   cudaFree(ctx->d_scalars);
   cudaFreeHost(ctx->h_scalars);
   for (cudaEvent_t e : ctx->ev) { cudaEventDestroy(e); }
   delete ctx;
}

int b2f_ctx_sync(b2f_ctx *ctx)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx, "b2f_ctx_sync: ctx is NULL");
   B2F_CUDA(cudaStreamSynchronize(ctx->stream));
   return B2F_OK;
}
int b2f_ctx_device(const b2f_ctx *ctx) { return ctx ? ctx->device : -1; }
int b2f_ctx_sm_count(const b2f_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
int64_t b2f_ctx_launch_count(const b2f_ctx *ctx) { return ctx ? ctx->launches : 0; }
void b2f_ctx_reset_launch_count(b2f_ctx *ctx) { if (ctx) { ctx->launches = 0; } }

int b2f_ctx_kernel_timing(b2f_ctx *ctx, int enable)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx, "b2f_ctx_kernel_timing: ctx is NULL");
   B2F_CUDA(cudaStreamSynchronize(ctx->stream));
   ctx->timing = enable != 0;
   ctx->ev_used = 0;
   return B2F_OK;
}

int b2f_ctx_kernel_time(b2f_ctx *ctx, double *ms_total_h, int *launches_h)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx && ms_total_h && launches_h, "b2f_ctx_kernel_time: NULL argument");
   B2F_CUDA(cudaStreamSynchronize(ctx->stream));
   double total = 0.0;
   const int n = ctx->ev_used / 2;
   for (int i = 0; i < n; i++)
   {
      float ms = 0.f;
      B2F_CUDA(cudaEventElapsedTime(&ms, ctx->ev[2 * i], ctx->ev[2 * i + 1]));
      total += ms;
   }
   *ms_total_h = total;
   *launches_h = n;
   ctx->ev_used = 0;
   return B2F_OK;
}

// ---- memory -----------------------------------------------------------------------------
int b2f_malloc(b2f_ctx *ctx, size_t bytes, void **dptr)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx && dptr, "b2f_malloc: NULL argument");
   B2F_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
   return B2F_OK;
}
int b2f_free(b2f_ctx *ctx, void *dptr)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx, "b2f_free: ctx is NULL");
   if (dptr) { B2F_CUDA(cudaFree(dptr)); }
   return B2F_OK;
}
int b2f_malloc_host(b2f_ctx *ctx, size_t bytes, void **hptr)
{
   B2F_REQUIRE(ctx && hptr, "b2f_malloc_host: NULL argument");
   B2F_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1));
   return B2F_OK;
}
int b2f_free_host(b2f_ctx *ctx, void *hptr)
{
   B2F_REQUIRE(ctx, "b2f_free_host: ctx is NULL");
   if (hptr) { B2F_CUDA(cudaFreeHost(hptr)); }
   return B2F_OK;
}
int b2f_memcpy_h2d(b2f_ctx *ctx, void *dst, const void *src_h, size_t bytes)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx, "b2f_memcpy_h2d: ctx is NULL");
   if (bytes) { B2F_CUDA(cudaMemcpyAsync(dst, src_h, bytes, cudaMemcpyHostToDevice, ctx->stream)); }
   return B2F_OK;
}
int b2f_memcpy_d2h(b2f_ctx *ctx, void *dst_h, const void *src, size_t bytes)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx, "b2f_memcpy_d2h: ctx is NULL");
   if (bytes) { B2F_CUDA(cudaMemcpyAsync(dst_h, src, bytes, cudaMemcpyDeviceToHost, ctx->stream)); }
   return B2F_OK;
}
int b2f_memcpy_d2d(b2f_ctx *ctx, void *dst, const void *src, size_t bytes)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx, "b2f_memcpy_d2d: ctx is NULL");
   if (bytes) { B2F_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream)); }
   return B2F_OK;
}
int b2f_memset(b2f_ctx *ctx, void *dst, int byte, size_t bytes)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx, "b2f_memset: ctx is NULL");
   if (bytes) { B2F_CUDA(cudaMemsetAsync(dst, byte, bytes, ctx->stream)); }
   return B2F_OK;
}

// ---- ElementRestriction -----------------------------------------------------------------
static int restriction_alloc(b2f_ctx *ctx, int ne, int dof, int ndofs, int vdim, int byvdim,
                             b2f_restriction **out)
{
   B2F_REQUIRE(ctx && out, "b2f_restriction_create: NULL argument");
   B2F_REQUIRE(ne >= 0 && dof >= 1 && ndofs >= 0 && vdim >= 1,
               "b2f_restriction_create: bad sizes ne=%d dof=%d ndofs=%d vdim=%d", ne, dof, ndofs, vdim);
   B2F_REQUIRE((int64_t)ne * dof <= INT32_MAX,
               "b2f_restriction_create: ne*dof = %lld exceeds the int32 E-index range",
               (long long)ne * dof);
   b2f_restriction *r = new b2f_restriction;
   r->ctx = ctx; r->ne = ne; r->dof = dof; r->ndofs = ndofs; r->vdim = vdim; r->byvdim = byvdim;
   const size_t n = (size_t)ne * dof;
   B2F_CUDA(cudaMalloc((void **)&r->gather_map, sizeof(int) * (n ? n : 1)));
   B2F_CUDA(cudaMalloc((void **)&r->indices, sizeof(int) * (n ? n : 1)));
   B2F_CUDA(cudaMalloc((void **)&r->offsets, sizeof(int) * ((size_t)ndofs + 1)));
   *out = r;
   return B2F_OK;
}

int b2f_restriction_create(b2f_ctx *ctx, int ne, int dof, int ndofs, int vdim, int byvdim,
                           const int *gather_map_h, const int *offsets_h, const int *indices_h,
                           b2f_restriction **out)
{
   B2F_ON(ctx);
   B2F_REQUIRE(gather_map_h && offsets_h && indices_h, "b2f_restriction_create: NULL table");
   b2f_restriction *r = nullptr;
   B2F_TRY(restriction_alloc(ctx, ne, dof, ndofs, vdim, byvdim, &r));
   const size_t n = (size_t)ne * dof;
   if (n)
   {
      B2F_CUDA(cudaMemcpy(r->gather_map, gather_map_h, sizeof(int) * n, cudaMemcpyHostToDevice));
      B2F_CUDA(cudaMemcpy(r->indices, indices_h, sizeof(int) * n, cudaMemcpyHostToDevice));
   }
   B2F_CUDA(cudaMemcpy(r->offsets, offsets_h, sizeof(int) * ((size_t)ndofs + 1),
                       cudaMemcpyHostToDevice));
   *out = r;
   return B2F_OK;
}

int b2f_restriction_create_from_elem_dof(b2f_ctx *ctx, int ne, int dof, int ndofs, int vdim,
                                         int byvdim, const int *elem_dof_h, const int *dof_map_h,
                                         b2f_restriction **out)
{
   B2F_ON(ctx);
   B2F_REQUIRE(elem_dof_h || ne == 0, "b2f_restriction_create_from_elem_dof: elem_dof_h is NULL");
   b2f_restriction *r = nullptr;
   B2F_TRY(restriction_alloc(ctx, ne, dof, ndofs, vdim, byvdim, &r));
   int *d_ed = nullptr, *d_dm = nullptr;
   B2F_TRY(upload(elem_dof_h, (size_t)ne * dof, &d_ed));
   if (dof_map_h) { B2F_TRY(upload(dof_map_h, (size_t)dof, &d_dm)); }
   int rc = restriction_build_device(r, d_ed, d_dm);
   cudaFree(d_ed);
   cudaFree(d_dm);
   if (rc) { b2f_restriction_destroy(r); return rc; }
   *out = r;
   return B2F_OK;
}

void b2f_restriction_destroy(b2f_restriction *r)
{
   if (!r) { return; }
   cudaFree(r->gather_map);
   cudaFree(r->offsets);
   cudaFree(r->indices);
   delete r;
}

int b2f_restriction_tables(const b2f_restriction *r, const int **gather_map, const int **offsets,
                           const int **indices)
{
   B2F_ON(r ? r->ctx : nullptr);
   B2F_REQUIRE(r, "b2f_restriction_tables: NULL handle");
   if (gather_map) { *gather_map = r->gather_map; }
   if (offsets) { *offsets = r->offsets; }
   if (indices) { *indices = r->indices; }
   return B2F_OK;
}

int b2f_restriction_mult(b2f_restriction *r, const double *x_l, double *y_e)
{
   B2F_ON(r ? r->ctx : nullptr);
   B2F_REQUIRE(r && (x_l || !r->ndofs) && (y_e || !r->ne), "b2f_restriction_mult: NULL argument");
   return restriction_gather(r, x_l, y_e, false);
}
int b2f_restriction_abs_mult(b2f_restriction *r, const double *x_l, double *y_e)
{
   B2F_ON(r ? r->ctx : nullptr);
   B2F_REQUIRE(r && (x_l || !r->ndofs) && (y_e || !r->ne), "b2f_restriction_abs_mult: NULL argument");
   return restriction_gather(r, x_l, y_e, true);
}
int b2f_restriction_mult_transpose(b2f_restriction *r, const double *x_e, double *y_l, int add)
{
   B2F_ON(r ? r->ctx : nullptr);
   B2F_REQUIRE(r && (x_e || !r->ne) && (y_l || !r->ndofs),
               "b2f_restriction_mult_transpose: NULL argument");
   return restriction_scatter(r, x_e, y_l, add ? 1 : 0);
}
int b2f_restriction_abs_mult_transpose(b2f_restriction *r, const double *x_e, double *y_l)
{
   B2F_ON(r ? r->ctx : nullptr);
   B2F_REQUIRE(r && (x_e || !r->ne) && (y_l || !r->ndofs),
               "b2f_restriction_abs_mult_transpose: NULL argument");
   return restriction_scatter(r, x_e, y_l, 2);
}

// ---- integrators ------------------------------------------------------------------------
static int pa_alloc(b2f_ctx *ctx, int kind, int dim, int d1d, int q1d, int64_t ne, int ncomp,
                    const double *B_h, const double *G_h, b2f_pa **out)
{
   B2F_REQUIRE(ctx && out && B_h, "pa setup: NULL argument");
   B2F_REQUIRE(dim == 2 || dim == 3, "pa setup: dim = %d (2 or 3 supported)", dim);
   B2F_REQUIRE(d1d >= 2 && q1d >= 1 && d1d * q1d <= 256, "pa setup: bad D1D=%d Q1D=%d", d1d, q1d);
   B2F_REQUIRE(ne >= 0, "pa setup: ne < 0");
   b2f_pa *pa = new b2f_pa;
   pa->ctx = ctx; pa->kind = kind; pa->dim = dim; pa->d1d = d1d; pa->q1d = q1d; pa->ne = ne;
   pa->ncomp = ncomp;
   pa->symmetric = !(kind == B2F_KIND_DIFFUSION && ncomp == dim * dim);
   memset(pa->B, 0, sizeof(pa->B));
   memset(pa->G, 0, sizeof(pa->G));
   memcpy(pa->B, B_h, sizeof(double) * d1d * q1d);
   if (G_h) { memcpy(pa->G, G_h, sizeof(double) * d1d * q1d); }
   // collocated (BP5): B is EXACTLY the identity (the reference's barycentric evaluation at a node
   // returns exactly 1 and 0, measured max|B - I| = 0) -> the interpolation is skipped.  Anything
   // else, however close, takes the general path.
   pa->collocated = 0;
   if (d1d == q1d)
   {
      double dev = 0.0;
      for (int q = 0; q < q1d; q++)
         for (int d = 0; d < d1d; d++)
         {
            const double v = pa->B[q + q1d * d] - (q == d ? 1.0 : 0.0);
            dev = fabs(v) > dev ? fabs(v) : dev;
         }
      pa->collocated = dev == 0.0;
   }
   int nq = 1;
   for (int i = 0; i < dim; i++) { nq *= q1d; }
   pa->data_size = (int64_t)nq * ncomp * ne;
   B2F_CUDA(cudaMalloc((void **)&pa->data, sizeof(double) * (pa->data_size ? pa->data_size : 1)));
   B2F_TRY(upload(pa->B, (size_t)d1d * q1d, &pa->d_B));
   B2F_TRY(upload(pa->G, (size_t)d1d * q1d, &pa->d_G));
   *out = pa;
   return B2F_OK;
}

int b2f_diffusion_setup(b2f_ctx *ctx, int dim, int d1d, int q1d, int64_t ne, const double *B_h,
                        const double *G_h, const double *W_h, const double *J,
                        const double *coeff, int coeff_dim, int64_t coeff_size, b2f_pa **out)
{
   B2F_ON(ctx);
   B2F_REQUIRE(G_h && W_h && (J || ne == 0) && coeff, "b2f_diffusion_setup: NULL argument");
   const int sym = dim * (dim + 1) / 2;
   B2F_REQUIRE(coeff_dim == 1 || coeff_dim == dim || coeff_dim == sym || coeff_dim == dim * dim,
               "b2f_diffusion_setup: coeff_dim = %d", coeff_dim);
   int nq = 1;
   for (int i = 0; i < dim; i++) { nq *= q1d; }
   const int coeff_const = coeff_size == coeff_dim;
   B2F_REQUIRE(coeff_const || coeff_size == (int64_t)coeff_dim * nq * ne,
               "b2f_diffusion_setup: coeff_size = %lld is neither %d (constant) nor %lld",
               (long long)coeff_size, coeff_dim, (long long)coeff_dim * nq * ne);
   const int ncomp = coeff_dim == dim * dim ? dim * dim : sym;
   b2f_pa *pa = nullptr;
   B2F_TRY(pa_alloc(ctx, B2F_KIND_DIFFUSION, dim, d1d, q1d, ne, ncomp, B_h, G_h, &pa));
   double *d_W = nullptr;
   B2F_TRY(upload(W_h, (size_t)nq, &d_W));
   int rc = diffusion_setup(ctx, dim, q1d, ne, d_W, J, coeff, coeff_dim, coeff_const, pa->data);
   cudaStreamSynchronize(ctx->stream);
   cudaFree(d_W);
   if (rc) { b2f_pa_destroy(pa); return rc; }
   *out = pa;
   return B2F_OK;
}

int b2f_mass_setup(b2f_ctx *ctx, int dim, int d1d, int q1d, int64_t ne, const double *B_h,
                   const double *W_h, const double *detJ, const double *coeff, int64_t coeff_size,
                   b2f_pa **out)
{
   B2F_ON(ctx);
   B2F_REQUIRE(W_h && (detJ || ne == 0) && coeff, "b2f_mass_setup: NULL argument");
   int nq = 1;
   for (int i = 0; i < dim; i++) { nq *= q1d; }
   const int coeff_const = coeff_size == 1;
   B2F_REQUIRE(coeff_const || coeff_size == (int64_t)nq * ne, "b2f_mass_setup: coeff_size = %lld",
               (long long)coeff_size);
   b2f_pa *pa = nullptr;
   B2F_TRY(pa_alloc(ctx, B2F_KIND_MASS, dim, d1d, q1d, ne, 1, B_h, nullptr, &pa));
   double *d_W = nullptr;
   B2F_TRY(upload(W_h, (size_t)nq, &d_W));
   int rc = mass_setup(ctx, nq, ne, d_W, detJ, coeff, coeff_const, pa->data);
   cudaStreamSynchronize(ctx->stream);
   cudaFree(d_W);
   if (rc) { b2f_pa_destroy(pa); return rc; }
   *out = pa;
   return B2F_OK;
}

int b2f_pa_setup_from_nodes(b2f_ctx *ctx, int kind, int dim, int d1d, int q1d, int64_t ne,
                            const double *B_h, const double *G_h, const double *W_h, int node_d1d,
                            const double *Bn_h, const double *Gn_h, const double *nodes_e,
                            double coeff, b2f_pa **out)
{
   B2F_ON(ctx);
   B2F_REQUIRE(kind == B2F_KIND_DIFFUSION || kind == B2F_KIND_MASS,
               "b2f_pa_setup_from_nodes: kind = %d", kind);
   B2F_REQUIRE(W_h && Bn_h && Gn_h && (nodes_e || ne == 0) && node_d1d >= 2,
               "b2f_pa_setup_from_nodes: NULL argument");
   B2F_REQUIRE(kind == B2F_KIND_MASS || G_h, "b2f_pa_setup_from_nodes: G_h is NULL");
   int nq = 1;
   for (int i = 0; i < dim; i++) { nq *= q1d; }
   const int ncomp = kind == B2F_KIND_MASS ? 1 : dim * (dim + 1) / 2;
   b2f_pa *pa = nullptr;
   B2F_TRY(pa_alloc(ctx, kind, dim, d1d, q1d, ne, ncomp, B_h, G_h, &pa));
   double *d_W = nullptr, *d_Bn = nullptr, *d_Gn = nullptr;
   B2F_TRY(upload(W_h, (size_t)nq, &d_W));
   B2F_TRY(upload(Bn_h, (size_t)q1d * node_d1d, &d_Bn));
   B2F_TRY(upload(Gn_h, (size_t)q1d * node_d1d, &d_Gn));
   int rc = setup_from_nodes(ctx, kind, dim, q1d, ne, d_W, node_d1d, d_Bn, d_Gn, nodes_e, coeff,
                             pa->data);
   cudaStreamSynchronize(ctx->stream);
   cudaFree(d_W); cudaFree(d_Bn); cudaFree(d_Gn);
   if (rc) { b2f_pa_destroy(pa); return rc; }
   *out = pa;
   return B2F_OK;
}

void b2f_pa_destroy(b2f_pa *pa)
{
   if (!pa) { return; }
   b2f_pa_destroy(pa->abs_twin);
   cudaFree(pa->data);
   cudaFree(pa->d_B);
   cudaFree(pa->d_G);
   delete pa;
}

int b2f_pa_data(const b2f_pa *pa, const double **data, int64_t *size)
{
   B2F_ON(pa ? pa->ctx : nullptr);
   B2F_REQUIRE(pa, "b2f_pa_data: NULL handle");
   if (data) { *data = pa->data; }
   if (size) { *size = pa->data_size; }
   return B2F_OK;
}

int b2f_pa_add_mult(b2f_pa *pa, const double *x_e, double *y_e)
{
   B2F_ON(pa ? pa->ctx : nullptr);
   B2F_REQUIRE(pa && (pa->ne == 0 || (x_e && y_e)), "b2f_pa_add_mult: NULL argument");
   return pa_apply(pa, x_e, y_e, nullptr);
}

} // extern "C"
__global__ void abs_copy_kernel(int64_t n, const double *__restrict__ src, double *__restrict__ dst)
{
   for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
   {
      dst[i] = fabs(src[i]);
   }
}
// the integrator with |B|, |G| and |pa_data| (DiffusionIntegrator::AddAbsMultPA,
// bilininteg_diffusion_pa.cpp:196-209; MassIntegrator::AddAbsMultPA, bilininteg_mass_pa.cpp:207-228)
static int pa_abs_twin(b2f_pa *pa, b2f_pa **out)
{
   if (!pa->abs_twin)
   {
      double Ba[256], Ga[256];
      for (int i = 0; i < 256; i++) { Ba[i] = fabs(pa->B[i]); Ga[i] = fabs(pa->G[i]); }
      b2f_pa *t = nullptr;
      B2F_TRY(pa_alloc(pa->ctx, pa->kind, pa->dim, pa->d1d, pa->q1d, pa->ne, pa->ncomp, Ba, Ga, &t));
      t->symmetric = pa->symmetric;
      if (pa->data_size)
      {
         abs_copy_kernel<<<(unsigned)std::min<int64_t>(cdiv(pa->data_size, 256), 148 * 64), 256, 0,
                           pa->ctx->stream>>>(pa->data_size, pa->data, t->data);
         if (cudaGetLastError() != cudaSuccess) { b2f_pa_destroy(t); B2F_REQUIRE(false, "abs twin: launch failed"); }
      }
      pa->abs_twin = t;
   }
   *out = pa->abs_twin;
   return B2F_OK;
}
extern "C" {

int b2f_pa_add_abs_mult(b2f_pa *pa, const double *x_e, double *y_e)
{
   B2F_ON(pa ? pa->ctx : nullptr);
   B2F_REQUIRE(pa && (pa->ne == 0 || (x_e && y_e)), "b2f_pa_add_abs_mult: NULL argument");
   b2f_pa *t = nullptr;
   B2F_TRY(pa_abs_twin(pa, &t));
   return pa_apply(t, x_e, y_e, nullptr);
}

int b2f_pa_add_mult_transpose(b2f_pa *pa, const double *x_e, double *y_e)
{
   B2F_ON(pa ? pa->ctx : nullptr);
   B2F_REQUIRE(pa && (pa->ne == 0 || (x_e && y_e)), "b2f_pa_add_mult_transpose: NULL argument");
   // bilininteg_diffusion_pa.cpp:95-106, bilininteg_mass_pa.cpp (AddMultTransposePA = AddMultPA)
   B2F_REQUIRE(pa->kind == B2F_KIND_MASS || pa->symmetric,
               "DiffusionIntegrator::AddMultTransposePA only implemented in the symmetric case.");
   return pa_apply(pa, x_e, y_e, nullptr);
}

int b2f_pa_add_mult_vec(b2f_pa *pa, int vdim, const double *x_e, double *y_e)
{
   B2F_ON(pa ? pa->ctx : nullptr);
   B2F_REQUIRE(pa && vdim >= 1 && (pa->ne == 0 || (x_e && y_e)), "b2f_pa_add_mult_vec: bad argument");
   if (vdim == 1) { return pa_apply(pa, x_e, y_e, nullptr); }
   pa->vd = vdim;
   const int rc = pa_apply(pa, x_e, y_e, nullptr);
   pa->vd = 1;
   return rc;
}

int b2f_domain_lf_assemble(b2f_ctx *ctx, int dim, int d1d, int q1d, int vdim, int64_t ne,
                           const int *markers, const double *B_h, const double *W_h, const double *detJ,
                           const double *coeff, int64_t coeff_size, double *y_e)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx && B_h && W_h && coeff && (ne == 0 || (detJ && y_e)), "b2f_domain_lf_assemble: NULL argument");
   B2F_REQUIRE(dim == 2 || dim == 3, "b2f_domain_lf_assemble: dim = %d", dim);
   B2F_REQUIRE(d1d >= 1 && d1d <= 16 && q1d >= 1 && q1d <= 16 && vdim >= 1,
               "b2f_domain_lf_assemble: D1D = %d, Q1D = %d, vdim = %d", d1d, q1d, vdim);
   int nq = 1;
   for (int i = 0; i < dim; i++) { nq *= q1d; }
   const int cst = coeff_size == vdim;
   B2F_REQUIRE(cst || coeff_size == (int64_t)vdim * nq * ne, "b2f_domain_lf_assemble: coeff_size = %lld",
               (long long)coeff_size);
   double *B_d = nullptr, *W_d = nullptr;
   B2F_TRY(upload(B_h, (size_t)d1d * q1d, &B_d));
   int rc = upload(W_h, (size_t)nq, &W_d);
   if (rc == B2F_OK) { rc = lf_domain(ctx, dim, d1d, q1d, vdim, ne, markers, B_d, W_d, detJ, coeff, cst, y_e); }
   if (rc == B2F_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = B2F_ERR_CUDA; }
   cudaFree(B_d); cudaFree(W_d);
   return rc;
}

int b2f_pa_assemble_diagonal(b2f_pa *pa, double *diag_e)
{
   B2F_ON(pa ? pa->ctx : nullptr);
   B2F_REQUIRE(pa && (pa->ne == 0 || diag_e), "b2f_pa_assemble_diagonal: NULL argument");
   return pa_diag(pa, diag_e);
}

// ---- operator ---------------------------------------------------------------------------
// out[c*n + i] = -1 where vdof (gmap[i], c) is essential, else gmap[i]; vdof = g*xs + c*xo
__global__ void mask_gather_map_kernel(int64_t n, int vd, int64_t xs, int64_t xo,
                                       const int *__restrict__ gmap,
                                       const unsigned char *__restrict__ is_ess, int *out)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   const int g = gmap[i];
   for (int c = 0; c < vd; c++) { out[c * n + i] = (g >= 0 && is_ess[g * xs + c * xo]) ? -1 : g; }
}

// y[vdof(g, c)] = x[g] for all components
__global__ void replicate_components_kernel(int nd, int vd, int64_t xs, int64_t xo,
                                            const double *__restrict__ x, double *y)
{
   const int g = blockIdx.x * blockDim.x + threadIdx.x;
   if (g >= nd) { return; }
   for (int c = 0; c < vd; c++) { y[g * xs + c * xo] = x[g]; }
}
__global__ void mark_kernel(int n, const int *__restrict__ idx, unsigned char *flags)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { flags[idx[i]] = 1; }
}
} // extern "C"
// dst row j = src row rows[j]   (row_len entries each)
template <typename T>
__global__ void compact_rows_kernel(int64_t n_rows, int64_t row_len, const int *__restrict__ rows,
                                    const T *__restrict__ src, T *__restrict__ dst)
{
   const int64_t total = n_rows * row_len;
   for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
        i += (int64_t)gridDim.x * blockDim.x)
   {
      const int64_t j = i / row_len, k = i - j * row_len;
      dst[i] = src[(int64_t)rows[j] * row_len + k];
   }
}
extern "C" {
// AddWithMarkers_ (bilinearform_ext.cpp:753-774): y_e += t_e on the marked elements
__global__ void add_marked_kernel(int64_t n, int64_t nd, const unsigned char *__restrict__ on,
                                  const double *__restrict__ t, double *y)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && on[i / nd]) { y[i] += t[i]; }
}
// assemble_diagonal_with_markers (bilinearform_ext.cpp:374-399): the accumulated E-diagonal
// of the unmarked elements is reset to 0
__global__ void zero_unmarked_kernel(int64_t n, int64_t nd, const unsigned char *__restrict__ on, double *d)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && !on[i / nd]) { d[i] = 0.0; }
}
__global__ void any_negative_kernel(int64_t n, const int *__restrict__ v, int *flag)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && v[i] < 0) { *flag = 1; }
}

static void marked_free_device(b2f_marked *m)
{
   if (!m) { return; }
   if (m->cpa) { cudaFree(m->cpa->data); delete m->cpa; m->cpa = nullptr; }
   cudaFree(m->idx_c); m->idx_c = nullptr;
   cudaFree(m->fidx_c); m->fidx_c = nullptr;
   cudaFree(m->on); m->on = nullptr;
}

// compact the enabled elements of integrator i (data slabs, gather rows, masked gather rows)
static int marked_rebuild(b2f_operator *op, size_t i)
{
   b2f_marked *m = op->marked[i];
   if (!m) { return B2F_OK; }
   b2f_ctx *ctx = op->ctx;
   const b2f_restriction *r = op->r;
   const b2f_pa *pa = op->integ[i];
   marked_free_device(m);
   std::vector<int> rows;
   for (int e = 0; e < r->ne; e++) { if (m->on_h[e]) { rows.push_back(e); } }
   m->ne_c = (int64_t)rows.size();
   B2F_TRY(upload(m->on_h.data(), m->on_h.size(), &m->on));
   int *d_rows = nullptr;
   B2F_TRY(upload(rows.data(), rows.size(), &d_rows));
   const int64_t slab = pa->ne ? pa->data_size / pa->ne : 0;
   m->cpa = new b2f_pa(*pa);
   m->cpa->abs_twin = nullptr;
   m->cpa->data = nullptr;
   m->cpa->ne = m->ne_c;
   m->cpa->data_size = slab * m->ne_c;
   const int64_t nidx = m->ne_c * r->dof;
   int rc = B2F_OK;
   auto run = [&]() -> int
   {
      B2F_CUDA(cudaMalloc((void **)&m->cpa->data, sizeof(double) * (size_t)(m->cpa->data_size + 2)));
      B2F_CUDA(cudaMalloc((void **)&m->idx_c, sizeof(int) * (size_t)(nidx + 4)));
      if (m->ne_c)
      {
         compact_rows_kernel<double><<<(unsigned)std::min<int64_t>(cdiv(m->cpa->data_size, 256), 65535 * 16), 256,
                                       0, ctx->stream>>>(m->ne_c, slab, d_rows, pa->data, m->cpa->data);
         B2F_LAUNCH_CHECK(ctx);
         compact_rows_kernel<int><<<(unsigned)std::min<int64_t>(cdiv(nidx, 256), 65535 * 16), 256, 0,
                                    ctx->stream>>>(m->ne_c, r->dof, d_rows, r->gather_map, m->idx_c);
         B2F_LAUNCH_CHECK(ctx);
      }
      if (op->fused_idx)
      {
         B2F_CUDA(cudaMalloc((void **)&m->fidx_c, sizeof(int) * (size_t)(nidx * r->vdim + 4)));
         for (int c = 0; c < r->vdim && m->ne_c; c++)
         {
            compact_rows_kernel<int><<<(unsigned)std::min<int64_t>(cdiv(nidx, 256), 65535 * 16), 256, 0,
                                       ctx->stream>>>(m->ne_c, r->dof, d_rows,
                                                      op->fused_idx + (int64_t)c * r->ne * r->dof,
                                                      m->fidx_c + c * nidx);
            B2F_LAUNCH_CHECK(ctx);
         }
      }
      B2F_CUDA(cudaStreamSynchronize(ctx->stream));
      return B2F_OK;
   };
   rc = run();
   cudaFree(d_rows);
   return rc;
}

// (re)build the essential-dof tables: ess_l = constrained LOCAL dofs (mask of the fused
// gather index); op->ess = the list ConstrainedOperator works on (ldofs, or tdofs when a
// halo is attached).
static int operator_build_masks(b2f_operator *op, const int *ess_ldofs_h, int n_ess_l)
{
   b2f_ctx *ctx = op->ctx;
   b2f_restriction *r = op->r;
   const int64_t n = (int64_t)r->ne * r->dof;
   cudaFree(op->fused_idx); op->fused_idx = nullptr;
   cudaFree(op->ess); op->ess = nullptr;
   op->n_ess = 0;
   std::vector<int> ess_c; // constrained list in the operator's own numbering
   if (op->halo)
   {
      const int *l2t = halo_ldof_to_tdof_h(op->halo);
      for (int i = 0; i < n_ess_l; i++)
      {
         const int t = l2t[ess_ldofs_h[i]];
         if (t >= 0) { ess_c.push_back(t); }
      }
   }
   else { ess_c.assign(ess_ldofs_h, ess_ldofs_h + n_ess_l); }
   op->n_ess = (int)ess_c.size();
   B2F_TRY(upload(ess_c.data(), ess_c.size(), &op->ess));
   if (n_ess_l > 0 && n > 0)
   {
      int *d_ess_l = nullptr;
      unsigned char *flags = nullptr;
      const int vd = r->vdim;
      const int64_t xs = r->byvdim ? vd : 1, xo = r->byvdim ? 1 : r->ndofs;
      const size_t nflags = (size_t)r->ndofs * vd + 1;
      B2F_TRY(upload(ess_ldofs_h, (size_t)n_ess_l, &d_ess_l));
      B2F_CUDA(cudaMalloc((void **)&flags, nflags));
      B2F_CUDA(cudaMemsetAsync(flags, 0, nflags, ctx->stream));
      mark_kernel<<<(unsigned)cdiv(n_ess_l, 256), 256, 0, ctx->stream>>>(n_ess_l, d_ess_l, flags);
      B2F_LAUNCH_CHECK(ctx);
      B2F_CUDA(cudaMalloc((void **)&op->fused_idx, sizeof(int) * n * vd));
      mask_gather_map_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(n, vd, xs, xo, r->gather_map,
                                                                              flags, op->fused_idx);
      B2F_LAUNCH_CHECK(ctx);
      B2F_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(flags);
      cudaFree(d_ess_l);
   }
   for (size_t i = 0; i < op->marked.size(); i++) { B2F_TRY(marked_rebuild(op, i)); }
   return B2F_OK;
}

int b2f_operator_create(b2f_ctx *ctx, b2f_restriction *r, b2f_pa *const *integrators,
                        int n_integrators, const int *ess_dofs_h, int n_ess, b2f_operator **out)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx && r && out && integrators && n_integrators >= 1,
               "b2f_operator_create: NULL argument");
   B2F_REQUIRE(n_ess == 0 || ess_dofs_h, "b2f_operator_create: ess_dofs_h is NULL");
   B2F_REQUIRE(r->vdim >= 1, "b2f_operator_create: vdim = %d", r->vdim);
   int dim_dof = 0;
   for (int i = 0; i < n_integrators; i++)
   {
      const b2f_pa *pa = integrators[i];
      B2F_REQUIRE(pa, "b2f_operator_create: integrator %d is NULL", i);
      int dof = 1;
      for (int k = 0; k < pa->dim; k++) { dof *= pa->d1d; }
      B2F_REQUIRE(dof == r->dof && pa->ne == r->ne,
                  "b2f_operator_create: integrator %d (D1D=%d dim=%d NE=%lld) does not match the "
                  "restriction (dof=%d NE=%d)", i, pa->d1d, pa->dim, (long long)pa->ne, r->dof, r->ne);
      dim_dof = dof;
   }
   (void)dim_dof;
   B2F_REQUIRE((int64_t)r->ndofs * r->vdim <= INT32_MAX,
               "b2f_operator_create: %lld vdofs exceed the int32 index range of the operator's dof lists",
               (long long)r->ndofs * r->vdim);
   for (int i = 0; i < n_ess; i++)
   {
      B2F_REQUIRE(ess_dofs_h[i] >= 0 && (int64_t)ess_dofs_h[i] < (int64_t)r->ndofs * r->vdim,
                  "b2f_operator_create: essential dof %d out of range", ess_dofs_h[i]);
   }
   b2f_operator *op = new b2f_operator;
   op->ctx = ctx;
   op->r = r;
   op->integ.assign(integrators, integrators + n_integrators);
   op->marked.assign((size_t)n_integrators, nullptr);
   op->ess_l_h.assign(ess_dofs_h, ess_dofs_h + n_ess);
   // signed gather maps (non-H1 spaces) cannot use the fused kernel's "-1 = skip" encoding
   {
      int *flag = nullptr;
      const int64_t n = (int64_t)r->ne * r->dof;
      B2F_CUDA(cudaMalloc((void **)&flag, sizeof(int)));
      B2F_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
      if (n)
      {
         any_negative_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(n, r->gather_map, flag);
         B2F_LAUNCH_CHECK(ctx);
      }
      int h = 0;
      B2F_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      B2F_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(flag);
      op->signed_map = h;
   }
   int rc = operator_build_masks(op, ess_dofs_h, n_ess);
   if (rc) { b2f_operator_destroy(op); return rc; }
   *out = op;
   return B2F_OK;
}

void b2f_operator_destroy(b2f_operator *op)
{
   if (!op) { return; }
   cudaFree(op->ess); cudaFree(op->fused_idx);
   cudaFree(op->xe); cudaFree(op->ye);
   cudaFree(op->z); cudaFree(op->w);
   cudaFree(op->xl); cudaFree(op->yl);
   for (double *p : op->pcg_vec) { cudaFree(p); }
   cudaFree(op->pcg_hist);
   cudaFree(op->dot_acc);
   cudaFree(op->te);
   for (b2f_marked *m : op->marked) { marked_free_device(m); delete m; }
   delete op;
}

// scratch vector of at least n doubles (grown on demand; capacities live in the operator)
static int ensure_cap(b2f_operator *op, double **p, int64_t n)
{
   int64_t &cap = op->scratch_cap[(const void *)p];
   if (*p && cap >= n) { return B2F_OK; }
   if (*p) { B2F_CUDA(cudaStreamSynchronize(op->ctx->stream)); cudaFree(*p); *p = nullptr; cap = 0; }
   B2F_CUDA(cudaMalloc((void **)p, sizeof(double) * (n ? n : 1)));
   cap = n;
   return B2F_OK;
}
#define ensure(p, n) ensure_cap(op, p, n)

// y_l = R^T (sum_i A_i) R x_l on LOCAL vectors; idx selects the (masked) gather index
static int local_mult_fused(b2f_operator *op, const double *x_l, double *y_l, const int *idx,
                            double *dot_acc = nullptr)
{
   b2f_ctx *ctx = op->ctx;
   const b2f_restriction *r = op->r;
   B2F_TRY(zero_vector(ctx, y_l, (int64_t)r->ndofs * r->vdim));
   for (size_t i = 0; i < op->integ.size(); i++)
   {
      b2f_pa *pa = op->integ[i];
      const int *ix = idx;
      int64_t ne_i = r->ne;
      if (const b2f_marked *m = op->marked[i])
      {
         // marked integrator: its compacted elements with their own rows of the same index
         if (m->ne_c == 0) { continue; }
         pa = m->cpa;
         ix = (idx == r->gather_map || !m->fidx_c) ? m->idx_c : m->fidx_c;
         ne_i = m->ne_c;
      }
      if (r->vdim > 1)
      {
         // Vector{Mass,Diffusion}Integrator: the scalar kernel per component, one quadrature-data
         // load for all of them; masked gather maps are per component, the plain one is shared
         pa->vd = r->vdim;
         pa->vxs = r->byvdim ? r->vdim : 1;
         pa->vxo = r->byvdim ? 1 : r->ndofs;
         pa->vics = idx == r->gather_map ? 0 : ne_i * r->dof;
      }
      const int rc = pa_apply(pa, x_l, y_l, ix, dot_acc);
      pa->vd = 1; pa->vxs = 1; pa->vxo = 0; pa->vics = 0;
      if (rc) { return rc; }
   }
   return B2F_OK;
}

// the elements [eb, ee) only, accumulating into y_l (no zero fill); even-odd kernels only
static int local_mult_range(b2f_operator *op, const double *x_l, double *y_l, const int *idx,
                            int64_t eb, int64_t ee, double *dot_acc)
{
   for (size_t i = 0; i < op->integ.size(); i++)
   {
      b2f_pa *pa = op->integ[i];
      const int *ix = idx;
      int64_t b = eb, e = ee;
      if (const b2f_marked *m = op->marked[i])
      {
         pa = m->cpa;
         ix = (idx == op->r->gather_map || !m->fidx_c) ? m->idx_c : m->fidx_c;
         b = m->prefix[eb]; e = m->prefix[ee]; // compaction keeps the element order
         if (b == e) { continue; }
      }
      pa->rb = b; pa->re = e;
      const int rc = pa_apply(pa, x_l, y_l, ix, dot_acc);
      pa->rb = 0; pa->re = -1;
      if (rc) { return rc; }
   }
   return B2F_OK;
}

static bool operator_can_split(const b2f_operator *op)
{
   if (op->signed_map || op->r->vdim > 1 || op->n_interior < 0 || getenv("B2F_NO_OVERLAP")) { return false; }
   for (const b2f_pa *pa : op->integ) { if (!pa_has_fused_dot(pa)) { return false; } }
   return true;
}

// every integrator can accumulate x.(Ax) inside its kernel (and the gather map is unsigned).
// Opt-in (B2F_FUSED_DOT=1): on B200 the solver loop runs at the 1 kW power cap, where the
// 16 B/dof pass it saves is paid back by the slightly heavier operator kernel (measured:
// 17.0 vs 17.0 GDOF/s for BP3 p=4), so the plain two-kernel form stays the default.
static bool operator_has_fused_dot(const b2f_operator *op)
{
   if (op->signed_map || op->r->vdim > 1 || op->diag_policy != 1 || !getenv("B2F_FUSED_DOT")) { return false; }
   for (const b2f_pa *pa : op->integ) { if (!pa_has_fused_dot(pa)) { return false; } }
   return true;
}

// use_abs: PABilinearFormExtension::AbsMult -> MultInternal(x, y, true): |R|, AddAbsMultPA, |R|^T
static int local_mult_unfused(b2f_operator *op, const double *x_l, double *y_l, bool use_abs = false)
{
   b2f_ctx *ctx = op->ctx;
   b2f_restriction *r = op->r;
   const int64_t ne_dof = (int64_t)r->ne * r->dof * r->vdim;
   B2F_TRY(ensure(&op->xe, ne_dof));
   B2F_TRY(ensure(&op->ye, ne_dof));
   B2F_TRY(restriction_gather(r, x_l, op->xe, use_abs));               // bilinearform_ext.cpp:537-544
   B2F_CUDA(cudaMemsetAsync(op->ye, 0, sizeof(double) * ne_dof, ctx->stream)); // :545
   for (size_t i = 0; i < op->integ.size(); i++)                       // :548
   {
      b2f_pa *pa = op->integ[i];
      if (use_abs) { B2F_TRY(pa_abs_twin(op->integ[i], &pa)); }
      double *out = op->ye;
      if (op->marked[i])
      {
         // AddMultWithMarkers (:807-833): into a zeroed temporary, then the masked add
         B2F_TRY(ensure(&op->te, ne_dof));
         B2F_CUDA(cudaMemsetAsync(op->te, 0, sizeof(double) * ne_dof, ctx->stream));
         out = op->te;
      }
      pa->vd = r->vdim;
      const int rc = pa_apply(pa, op->xe, out, nullptr);
      pa->vd = 1;
      if (rc) { return rc; }
      if (op->marked[i] && ne_dof)
      {
         add_marked_kernel<<<(unsigned)cdiv(ne_dof, 256), 256, 0, ctx->stream>>>(
            ne_dof, (int64_t)r->dof * r->vdim, op->marked[i]->on, op->te, op->ye);
         B2F_LAUNCH_CHECK(ctx);
      }
   }
   return restriction_scatter(r, op->ye, y_l, use_abs ? 2 : 0);        // :551-558
}

static int local_mult(b2f_operator *op, const double *x_l, double *y_l, bool constrained_mask,
                      double *dot_acc = nullptr)
{
   if (op->signed_map)
   {
      // non-H1 signs: the reference sequence; masking is done by the caller's z = x, z[ess] = 0
      return local_mult_unfused(op, x_l, y_l);
   }
   const int *idx = (constrained_mask && op->fused_idx) ? op->fused_idx : op->r->gather_map;
   return local_mult_fused(op, x_l, y_l, idx, dot_acc);
}

// A on the operator's own vectors (true-dof vectors when a halo is attached):
// RAPOperator::Mult (linalg/operator.hpp:998): P, A, P^T.
// lcap: x and y are caller-owned buffers with room for n_ldofs entries (the PCG work
// vectors); with the owned-first numbering P and P^T then run in place and no vector is
// copied.  Otherwise the true-dof vectors go through the L-vector scratch xl / yl.
static int rap_mult(b2f_operator *op, const double *x, double *y, bool constrained_mask, bool lcap = false,
                    double *dot_acc = nullptr)
{
   if (!op->halo) { return local_mult(op, x, y, constrained_mask, dot_acc); }
   b2f_ctx *ctx = op->ctx;
   const int64_t nl = (int64_t)op->r->ndofs * op->r->vdim; // local (v)dofs; the halo tables are per vdof
   if (halo_is_prefix(op->halo))
   {
      int nl2 = 0, nt = 0;
      halo_sizes(op->halo, &nl2, &nt);
      double *xl = const_cast<double *>(x), *yl = y;
      if (!lcap)
      {
         B2F_TRY(ensure(&op->xl, nl));
         B2F_TRY(ensure(&op->yl, nl));
         xl = op->xl; yl = op->yl;
         B2F_TRY(copy_vector(ctx, xl, x, nt));
      }
      if (halo_is_p2p(op->halo) && operator_can_split(op))
      {
         // The exchange is one-sided (peer stores + epoch flags), so on ONE stream it overlaps
         // with element work by itself: push my shared values, run the first half of the
         // interior elements (they read no neighbour-owned dof) while the neighbours' values
         // arrive, unpack, run the boundary elements, push their contributions back, run the
         // second half of the interior while those travel, then assemble what arrived.
         const int *idx = (constrained_mask && op->fused_idx) ? op->fused_idx : op->r->gather_map;
         const int64_t ni = op->n_interior, ne = op->r->ne, half = ni / 2;
         B2F_TRY(halo_p_push(op->halo, xl));
         B2F_TRY(zero_vector(ctx, yl, nl));
         B2F_TRY(local_mult_range(op, xl, yl, idx, 0, half, dot_acc));
         B2F_TRY(halo_p_wait(op->halo, xl));
         B2F_TRY(local_mult_range(op, xl, yl, idx, ni, ne, dot_acc));
         B2F_TRY(halo_pt_push(op->halo, yl));
         B2F_TRY(local_mult_range(op, xl, yl, idx, half, ni, dot_acc));
         B2F_TRY(halo_pt_wait(op->halo, yl));
      }
      else
      {
         B2F_TRY(halo_mult_inplace(op->halo, xl));
         B2F_TRY(local_mult(op, xl, yl, constrained_mask, dot_acc));
         B2F_TRY(halo_mult_transpose_inplace(op->halo, yl));
      }
      if (!lcap) { B2F_TRY(copy_vector(ctx, y, yl, nt)); }
      return B2F_OK;
   }
   B2F_TRY(ensure(&op->xl, nl));
   B2F_TRY(ensure(&op->yl, nl));
   B2F_TRY(halo_mult(op->halo, x, op->xl));
   B2F_TRY(local_mult(op, op->xl, op->yl, constrained_mask, dot_acc));
   return halo_mult_transpose(op->halo, op->yl, y);
}

int b2f_operator_set_markers(b2f_operator *op, int integrator, const int *elem_attributes_h,
                             const int *markers_h, int n_markers)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op, "b2f_operator_set_markers: NULL handle");
   B2F_REQUIRE(integrator >= 0 && integrator < (int)op->integ.size(),
               "b2f_operator_set_markers: integrator %d of %d", integrator, (int)op->integ.size());
   if (b2f_marked *old = op->marked[integrator])
   {
      B2F_CUDA(cudaStreamSynchronize(op->ctx->stream));
      marked_free_device(old);
      delete old;
      op->marked[integrator] = nullptr;
   }
   if (!markers_h) { return B2F_OK; } // acts on every element again
   const int ne = op->r->ne;
   B2F_REQUIRE(elem_attributes_h || ne == 0, "b2f_operator_set_markers: elem_attributes_h is NULL");
   b2f_marked *m = new b2f_marked;
   m->on_h.resize((size_t)ne);
   m->prefix.resize((size_t)ne + 1);
   m->prefix[0] = 0;
   for (int e = 0; e < ne; e++)
   {
      const int attr = elem_attributes_h[e];
      if (attr > n_markers)
      {
         delete m;
         B2F_REQUIRE(false, "b2f_operator_set_markers: element %d has attribute %d, the marker array %d entries",
                     e, attr, n_markers);
      }
      m->on_h[e] = !(attr <= 0 || markers_h[attr - 1] == 0); // bilinearform_ext.cpp:768
      m->prefix[e + 1] = m->prefix[e] + m->on_h[e];
   }
   op->marked[integrator] = m;
   return marked_rebuild(op, (size_t)integrator);
}

int b2f_operator_set_diag_policy(b2f_operator *op, int policy)
{
   B2F_REQUIRE(op && policy >= 0 && policy <= 2, "b2f_operator_set_diag_policy: bad argument");
   op->diag_policy = policy;
   return B2F_OK;
}

int b2f_operator_assemble_diagonal_constrained(b2f_operator *op, double *diag)
{
   B2F_ON(op ? op->ctx : nullptr);
   // ConstrainedOperator::AssembleDiagonal (linalg/operator.cpp:543-572)
   B2F_TRY(b2f_operator_assemble_diagonal(op, diag));
   if (op->diag_policy == 2) { return B2F_OK; }
   return set_entries(op->ctx, diag, op->ess, op->n_ess, op->diag_policy == 1 ? 1.0 : 0.0);
}

int b2f_operator_set_interior(b2f_operator *op, int64_t n_interior)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op, "b2f_operator_set_interior: NULL handle");
   B2F_REQUIRE(n_interior >= -1 && n_interior <= op->r->ne, "b2f_operator_set_interior: %lld of %d elements",
               (long long)n_interior, op->r->ne);
   op->n_interior = n_interior;
   return B2F_OK;
}

int b2f_operator_height(const b2f_operator *op)
{
   if (!op) { return -1; }
   if (op->halo)
   {
      int nl = 0, nt = 0;
      halo_sizes(op->halo, &nl, &nt);
      return nt;
   }
   return op->r->ndofs * op->r->vdim;
}

int b2f_operator_mult(b2f_operator *op, const double *x_l, double *y_l)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op && x_l && y_l, "b2f_operator_mult: NULL argument");
   B2F_REQUIRE(x_l != y_l, "b2f_operator_mult: x and y alias");
   return rap_mult(op, x_l, y_l, false);
}

int b2f_operator_mult_unfused(b2f_operator *op, const double *x_l, double *y_l)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op && x_l && y_l, "b2f_operator_mult_unfused: NULL argument");
   B2F_REQUIRE(!op->halo, "b2f_operator_mult_unfused: local operators only");
   return local_mult_unfused(op, x_l, y_l);
}

// y[ess] after the constrained product, by Operator::DiagonalPolicy (operator.cpp:631-658)
static int constrained_entries(b2f_operator *op, double *y, const double *x)
{
   if (op->diag_policy == 1) { return copy_entries(op->ctx, y, x, op->ess, op->n_ess); }
   if (op->diag_policy == 0) { return set_entries(op->ctx, y, op->ess, op->n_ess, 0.0); }
   B2F_REQUIRE(op->n_ess == 0, "ConstrainedOperator::Mult: DIAG_KEEP needs the action of the operator diagonal");
   return B2F_OK;
}

// Operator::AbsMult (linalg/operator.hpp:96) of the PA operator: RAPOperator::AbsMult
// (operator.hpp:1002; the conforming P has 0/1 entries, so |P| = P) around
// PABilinearFormExtension::AbsMult = MultInternal(x, y, useAbs = true) (bilinearform_ext.cpp:487-564)
static int rap_abs_mult(b2f_operator *op, const double *x, double *y)
{
   B2F_REQUIRE(op->r->vdim == 1, "AbsMult: scalar spaces only");
   if (!op->halo) { return local_mult_unfused(op, x, y, true); }
   const int nl = op->r->ndofs;
   B2F_TRY(ensure(&op->xl, nl));
   B2F_TRY(ensure(&op->yl, nl));
   B2F_TRY(halo_mult(op->halo, x, op->xl));
   B2F_TRY(local_mult_unfused(op, op->xl, op->yl, true));
   return halo_mult_transpose(op->halo, op->yl, y);
}

int b2f_operator_abs_mult(b2f_operator *op, const double *x, double *y)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op && x && y, "b2f_operator_abs_mult: NULL argument");
   B2F_REQUIRE(x != y, "b2f_operator_abs_mult: x and y alias");
   return rap_abs_mult(op, x, y);
}

int b2f_operator_abs_mult_constrained(b2f_operator *op, const double *x, double *y)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op && x && y, "b2f_operator_abs_mult_constrained: NULL argument");
   B2F_REQUIRE(x != y, "b2f_operator_abs_mult_constrained: x and y alias");
   if (op->n_ess == 0) { return rap_abs_mult(op, x, y); }
   // ConstrainedOperator::ConstrainedAbsMult (linalg/operator.cpp:663-722, DIAG_ONE)
   b2f_ctx *ctx = op->ctx;
   const int n = b2f_operator_height(op);
   B2F_TRY(ensure(&op->z, n));
   B2F_CUDA(cudaMemcpyAsync(op->z, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
   B2F_TRY(set_entries(ctx, op->z, op->ess, op->n_ess, 0.0));
   B2F_TRY(rap_abs_mult(op, op->z, y));
   return constrained_entries(op, y, x);
}

// dot_acc != nullptr: additionally leave (x, y) = sum_e z_e.(A_e z_e) + sum_ess x_i^2 in dot_acc
// (local part; z = x with the constrained entries zeroed)
static int mult_constrained(b2f_operator *op, const double *x, double *y, bool lcap, double *dot_acc = nullptr)
{
   b2f_ctx *ctx = op->ctx;
   if (dot_acc)
   {
      B2F_TRY(rap_mult(op, x, y, true, lcap, dot_acc));
      return copy_entries_sqsum(ctx, y, x, op->ess, op->n_ess, dot_acc);
   }
   if (op->signed_map && op->n_ess)
   {
      // operator.cpp:601-660 literally: z = x; z[ess] = 0; y = A z; y[ess] = x[ess]
      const int n = b2f_operator_height(op);
      B2F_TRY(ensure(&op->z, n));
      B2F_CUDA(cudaMemcpyAsync(op->z, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
      B2F_TRY(set_entries(ctx, op->z, op->ess, op->n_ess, 0.0));
      B2F_TRY(rap_mult(op, op->z, y, false));
   }
   else
   {
      // same result without the copy: constrained dofs read as 0 in the gather and are
      // skipped by the scatter
      B2F_TRY(rap_mult(op, x, y, true, lcap));
   }
   return constrained_entries(op, y, x);
}

int b2f_operator_mult_constrained(b2f_operator *op, const double *x, double *y)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op && x && y, "b2f_operator_mult_constrained: NULL argument");
   B2F_REQUIRE(x != y, "b2f_operator_mult_constrained: x and y alias");
   return mult_constrained(op, x, y, false);
}

int b2f_operator_assemble_diagonal(b2f_operator *op, double *diag)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op && diag, "b2f_operator_assemble_diagonal: NULL argument");
   b2f_ctx *ctx = op->ctx;
   b2f_restriction *r = op->r;
   const int64_t ne_dof = (int64_t)r->ne * r->dof;
   bool any_marked = false;
   for (size_t i = 0; i < op->integ.size(); i++) { any_marked = any_marked || op->marked[i]; }
   if (r->vdim == 1 && !any_marked)
   {
      // hot form: every integrator's element diagonal goes straight into the L-vector with unsigned
      // reductions (AssembleDiagonalPA + AbsMultTranspose, bilinearform_ext.cpp:370-430, in one
      // kernel per integrator); the E-vector diagonal never exists
      double *dl = diag;
      if (op->halo)
      {
         B2F_TRY(ensure(&op->yl, r->ndofs));
         dl = op->yl;
      }
      B2F_CUDA(cudaMemsetAsync(dl, 0, sizeof(double) * (size_t)r->ndofs, ctx->stream));
      for (size_t i = 0; i < op->integ.size(); i++) { B2F_TRY(pa_diag_fused(op->integ[i], r->gather_map, dl)); }
      // ParBilinearForm::AssembleDiagonal (fem/pbilinearform.cpp:293): P^T of the local diagonal
      return op->halo ? halo_mult_transpose(op->halo, op->yl, diag) : B2F_OK;
   }
   B2F_TRY(ensure(&op->ye, ne_dof));
   B2F_CUDA(cudaMemsetAsync(op->ye, 0, sizeof(double) * ne_dof, ctx->stream)); // bilinearform_ext.cpp:385
   for (size_t i = 0; i < op->integ.size(); i++)
   {
      B2F_TRY(pa_diag(op->integ[i], op->ye));
      if (op->marked[i] && ne_dof)
      {
         // the reference resets what has been accumulated so far on the unmarked elements
         zero_unmarked_kernel<<<(unsigned)cdiv(ne_dof, 256), 256, 0, ctx->stream>>>(ne_dof, r->dof,
                                                                                   op->marked[i]->on, op->ye);
         B2F_LAUNCH_CHECK(ctx);
      }
   }
   if (r->vdim > 1)
   {
      // vector integrators: the scalar diagonal for every component
      // (bilininteg_vecdiffusion_pa.cpp / bilininteg_vecmass_pa.cpp AssembleDiagonalPA)
      B2F_TRY(ensure(&op->z, (int64_t)r->ndofs * r->vdim)); // (z is also the vdof-sized scratch elsewhere)
      b2f_restriction scalar = *r;
      scalar.vdim = 1; scalar.byvdim = 0;
      B2F_TRY(restriction_scatter(&scalar, op->ye, op->z, 2));
      double *dl = diag;
      if (op->halo)
      {
         // ParBilinearForm::AssembleDiagonal (fem/pbilinearform.cpp:293): P^T of the local diagonal
         B2F_TRY(ensure(&op->yl, (int64_t)r->ndofs * r->vdim));
         dl = op->yl;
      }
      replicate_components_kernel<<<(unsigned)cdiv(r->ndofs, 256), 256, 0, ctx->stream>>>(
         r->ndofs, r->vdim, r->byvdim ? r->vdim : 1, r->byvdim ? 1 : r->ndofs, op->z, dl);
      B2F_LAUNCH_CHECK(ctx);
      return op->halo ? halo_mult_transpose(op->halo, op->yl, diag) : B2F_OK;
   }
   if (!op->halo) { return restriction_scatter(r, op->ye, diag, 2); } // AbsMultTranspose :426
   // ParBilinearForm::AssembleDiagonal (fem/pbilinearform.cpp:293): P^T of the local diagonal
   B2F_TRY(ensure(&op->yl, r->ndofs));
   B2F_TRY(restriction_scatter(r, op->ye, op->yl, 2));
   return halo_mult_transpose(op->halo, op->yl, diag);
}

int b2f_operator_eliminate_rhs(b2f_operator *op, const double *x, double *b)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op && x && b, "b2f_operator_eliminate_rhs: NULL argument");
   b2f_ctx *ctx = op->ctx;
   const int n = b2f_operator_height(op);
   if (op->n_ess == 0) { return B2F_OK; }
   // operator.cpp:574-599: w = 0; w[ess] = x[ess]; z = A w; b -= z; b[ess] = x[ess]
   B2F_TRY(ensure(&op->w, n));
   B2F_TRY(ensure(&op->z, n));
   B2F_CUDA(cudaMemsetAsync(op->w, 0, sizeof(double) * n, ctx->stream));
   B2F_TRY(copy_entries(ctx, op->w, x, op->ess, op->n_ess));
   B2F_TRY(rap_mult(op, op->w, op->z, false));
   B2F_TRY(vector_add(ctx, n, b, -1.0, op->z, b));
   return copy_entries(ctx, b, x, op->ess, op->n_ess);
}

int b2f_operator_set_halo(b2f_operator *op, b2f_halo *halo)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op, "b2f_operator_set_halo: NULL handle");
   if (halo)
   {
      int nl = 0, nt = 0;
      halo_sizes(halo, &nl, &nt);
      // vector spaces: the halo is built on vdofs (every table entry expanded per component by
      // the caller, as ParFiniteElementSpace's group communicator does for vdim > 1); with
      // Ordering::byVDIM the owned-first numbering keeps true vdofs the head of the local vector
      B2F_REQUIRE(op->r->vdim == 1 || op->r->byvdim,
                  "b2f_operator_set_halo: rank-parallel vector spaces need Ordering::byVDIM");
      B2F_REQUIRE(nl == op->r->ndofs * op->r->vdim, "b2f_operator_set_halo: halo has %d ldofs, the space %d",
                  nl, op->r->ndofs * op->r->vdim);
   }
   op->halo = halo;
   cudaFree(op->z); op->z = nullptr;
   cudaFree(op->w); op->w = nullptr;
   for (double *&p : op->pcg_vec) { cudaFree(p); p = nullptr; }
   return operator_build_masks(op, op->ess_l_h.data(), (int)op->ess_l_h.size());
}

// ---- Jacobi + BLAS-1 --------------------------------------------------------------------
int b2f_jacobi_setup(b2f_ctx *ctx, int n, const double *diag, const int *ess_dofs, int n_ess,
                     double damping, double *dinv)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx && (n == 0 || (diag && dinv)) && (n_ess == 0 || ess_dofs),
               "b2f_jacobi_setup: NULL argument");
   return jacobi_setup(ctx, n, diag, ess_dofs, n_ess, damping, dinv);
}

int b2f_jacobi_mult(b2f_ctx *ctx, int n, const double *dinv, const double *x, double *y)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx && (n == 0 || (dinv && x && y)), "b2f_jacobi_mult: NULL argument");
   return vector_mul(ctx, n, dinv, x, y);
}

int b2f_vector_dot(b2f_ctx *ctx, int64_t n, const double *x, const double *y, double *result_h)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx && result_h && (n == 0 || (x && y)), "b2f_vector_dot: NULL argument");
   if (n == 0) { *result_h = 0.0; return B2F_OK; }
   B2F_TRY(dot_to_device(ctx, n, x, y, ctx->d_scalars + 63));
   B2F_TRY(publish_scalars(ctx, ctx->d_scalars + 63, ctx->h_scalars + 63, 1));
   *result_h = ctx->h_scalars[63];
   return B2F_OK;
}

int b2f_vector_add(b2f_ctx *ctx, int64_t n, const double *x, double a, const double *y, double *z)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx && (n == 0 || (x && y && z)), "b2f_vector_add: NULL argument");
   return vector_add(ctx, n, x, a, y, z);
}

// ---- OperatorChebyshevSmoother (linalg/solvers.cpp:453-658) ------------------------------
} // extern "C"

struct b2f_chebyshev
{
   b2f_operator *op;
   int order;
   double coeffs[5];
   double *dinv = nullptr;
   double *res = nullptr, *help = nullptr; // "residual" and "helperVector", L-vector capacity
   double *vin = nullptr;                  // staging copy when the caller's input cannot be used in place
};

// z = p(D^-1 A) D^-1 r
static int chebyshev_apply(b2f_chebyshev *c, const double *x, double *y)
{
   b2f_operator *op = c->op;
   b2f_ctx *ctx = op->ctx;
   const int64_t n = b2f_operator_height(op);
   B2F_TRY(cheb_step(ctx, n, x, c->dinv, c->coeffs[0], c->res, y, true));
   for (int k = 1; k < c->order; k++)
   {
      B2F_TRY(mult_constrained(op, c->res, c->help, true));
      B2F_TRY(cheb_step(ctx, n, c->help, c->dinv, c->coeffs[k], c->res, y, false));
   }
   return B2F_OK;
}

extern "C" {

int b2f_chebyshev_coefficients_h(int order, double max_eig_estimate, double *coeffs_h)
{
   B2F_REQUIRE(coeffs_h, "b2f_chebyshev_coefficients_h: NULL argument");
   B2F_REQUIRE(order >= 1 && order <= 5, "Chebyshev smoother not implemented for order = %d", order);
   // Eigenvalue window [0.3, 1.2] lambda_max of D^-1 A (solvers.cpp:553-560); the coefficients
   // are those of the polynomial q with 1 - t q(t) = T_k((theta - t)/delta) / T_k(theta/delta),
   // written in the monomial basis of t; closed forms as in solvers.cpp:562-618.
   const double ub = 1.2 * max_eig_estimate, lb = 0.3 * max_eig_estimate;
   const double th = 0.5 * (ub + lb), de = 0.5 * (ub - lb);
   const double th2 = pow(th, 2), de2 = pow(de, 2);
   double *c = coeffs_h;
   switch (order)
   {
      case 1:
         c[0] = 1.0 / th;
         break;
      case 2:
      {
         const double w = 1.0 / (de2 - 2 * th2);
         c[0] = -4 * th * w;
         c[1] = 2 * w;
         break;
      }
      case 3:
      {
         const double a = 3 * de2;
         const double w = 1.0 / (-4 * pow(th, 3) + th * a);
         c[0] = w * (a - 12 * th2);
         c[1] = 12 / (a - 4 * th2);
         c[2] = -4 * w;
         break;
      }
      case 4:
      {
         const double a = 8 * de2;
         const double w = 1.0 / (pow(de, 4) + 8 * pow(th, 4) - th2 * a);
         c[0] = w * (32 * pow(th, 3) - 16 * th * de2);
         c[1] = w * (-48 * th2 + a);
         c[2] = 32 * th * w;
         c[3] = -8 * w;
         break;
      }
      case 5:
      {
         const double a = 5 * pow(de, 4), th4 = pow(th, 4);
         const double b60 = 60 * de2, b20 = 20 * de2;
         const double w = 1.0 / (16 * pow(th, 5) - pow(th, 3) * b20 + th * a);
         const double t160 = 160 * th2;
         const double v = 1.0 / (a + 16 * th4 - th2 * b20);
         c[0] = w * (a + 80 * th4 - th2 * b60);
         c[1] = v * (b60 - t160);
         c[2] = w * (-b20 + t160);
         c[3] = -80 * v;
         c[4] = 16 * w;
         break;
      }
   }
   return B2F_OK;
}

int b2f_power_method(b2f_operator *op, const double *dinv, double *v0, int num_steps, double tolerance,
                     double *eigenvalue_h)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op && dinv && v0 && eigenvalue_h, "b2f_power_method: NULL argument");
   b2f_ctx *ctx = op->ctx;
   b2f_halo *halo = op->halo;
   const int64_t n = b2f_operator_height(op);
   const int64_t ncap = halo ? (int64_t)op->r->ndofs * op->r->vdim : n;
   // the iteration vector and its image live in the PCG scratch (L-vector capacity)
   for (int k = 0; k < 3; k++) { B2F_TRY(ensure(&op->pcg_vec[k], ncap)); }
   double *va = op->pcg_vec[0], *vb = op->pcg_vec[1], *q = op->pcg_vec[2];
   double *slot = ctx->d_scalars + 62;
   B2F_CUDA(cudaMemcpyAsync(va, v0, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
   double eig = 1.0;
   for (int it = 0; it < num_steps; it++)
   {
      B2F_TRY(dot_to_device(ctx, n, va, va, slot));
      if (halo) { B2F_TRY(halo_allreduce_sum(halo, slot, 1)); }
      B2F_TRY(normalize_by(ctx, n, slot, va));                 // v0 /= sqrt(normV0)
      B2F_TRY(mult_constrained(op, va, q, true));              // ProductOperator: A first,
      B2F_TRY(vector_mul(ctx, n, dinv, q, vb));                // then the Jacobi smoother
      B2F_TRY(dot_to_device(ctx, n, va, vb, slot));
      if (halo) { B2F_TRY(halo_allreduce_sum(halo, slot, 1)); }
      B2F_TRY(publish_scalars(ctx, slot, ctx->h_scalars + 62, 1));
      const double eig_new = ctx->h_scalars[62];
      const double diff = fabs((eig_new - eig) / eig);
      eig = eig_new;
      std::swap(va, vb);
      if (diff < tolerance) { break; }
   }
   B2F_CUDA(cudaMemcpyAsync(v0, va, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
   B2F_CUDA(cudaStreamSynchronize(ctx->stream));
   *eigenvalue_h = eig;
   return B2F_OK;
}

int b2f_chebyshev_create(b2f_operator *op, const double *diag, int order, double max_eig_estimate,
                         b2f_chebyshev **out)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(op && diag && out, "b2f_chebyshev_create: NULL argument");
   *out = nullptr;
   double coeffs[5] = {0, 0, 0, 0, 0};
   B2F_TRY(b2f_chebyshev_coefficients_h(order, max_eig_estimate, coeffs));
   b2f_chebyshev *c = new b2f_chebyshev;
   c->op = op;
   c->order = order;
   memcpy(c->coeffs, coeffs, sizeof(coeffs));
   const int64_t n = b2f_operator_height(op);
   const int64_t ncap = op->halo ? (int64_t)op->r->ndofs * op->r->vdim : n;
   cudaError_t e1 = cudaMalloc((void **)&c->dinv, sizeof(double) * (n ? n : 1));
   cudaError_t e2 = cudaMalloc((void **)&c->res, sizeof(double) * (ncap ? ncap : 1));
   cudaError_t e3 = cudaMalloc((void **)&c->help, sizeof(double) * (ncap ? ncap : 1));
   if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
   {
      b2f_chebyshev_destroy(c);
      B2F_REQUIRE(false, "b2f_chebyshev_create: out of device memory");
   }
   // Setup(): dinv = 1/diag, dinv[ess] = 1 (solvers.cpp:538-550)
   int rc = jacobi_setup(op->ctx, (int)n, diag, op->ess, op->n_ess, 1.0, c->dinv);
   if (rc != B2F_OK) { b2f_chebyshev_destroy(c); return rc; }
   *out = c;
   return B2F_OK;
}

void b2f_chebyshev_destroy(b2f_chebyshev *c)
{
   if (!c) { return; }
   cudaFree(c->dinv); cudaFree(c->res); cudaFree(c->help); cudaFree(c->vin);
   delete c;
}

int b2f_chebyshev_mult(b2f_chebyshev *c, const double *x, double *y)
{
   B2F_ON((c && c->op) ? c->op->ctx : nullptr);
   B2F_REQUIRE(c && x && y, "b2f_chebyshev_mult: NULL argument");
   B2F_REQUIRE(x != y, "b2f_chebyshev_mult: x and y alias");
   return chebyshev_apply(c, x, y);
}

// ---- PCG --------------------------------------------------------------------------------
} // extern "C"

static int pcg_impl(b2f_operator *op, const double *dinv, b2f_chebyshev *cheb, const double *b, double *x,
                    const b2f_pcg_options *opt, b2f_pcg_result *result_h, double *history_h,
                    int history_capacity)
{
   B2F_REQUIRE(op && b && x && opt && result_h, "b2f_pcg: NULL argument");
   B2F_REQUIRE(opt->max_iter >= 0, "b2f_pcg: max_iter < 0");
   B2F_REQUIRE(pcg_state_bytes() <= (int)(48 * sizeof(double)), "b2f_pcg: state too large");
   b2f_ctx *ctx = op->ctx;
   b2f_halo *halo = op->halo;
   const int64_t n = b2f_operator_height(op);
   const int check_every = opt->check_every > 0 ? opt->check_every : 1;
   void *st = ctx->d_scalars; // PcgState lives in the context's device scalars
   // d and q (the operator's input and output) get room for a local-dof vector so that the
   // halo can run in place on them
   const int64_t ncap = halo ? (int64_t)op->r->ndofs * op->r->vdim : n;
   for (int k = 0; k < 4; k++) { B2F_TRY(ensure(&op->pcg_vec[k], ncap)); }
   double *r = op->pcg_vec[0], *d = op->pcg_vec[1], *z = op->pcg_vec[2], *q = op->pcg_vec[3];
   double *hist = nullptr;
   if (history_h && history_capacity > 0)
   {
      if (op->pcg_hist_cap < history_capacity)
      {
         cudaFree(op->pcg_hist);
         op->pcg_hist = nullptr;
         B2F_CUDA(cudaMalloc((void **)&op->pcg_hist, sizeof(double) * history_capacity));
         op->pcg_hist_cap = history_capacity;
      }
      hist = op->pcg_hist;
      B2F_CUDA(cudaMemsetAsync(hist, 0, sizeof(double) * history_capacity, ctx->stream));
   }
   auto reduce = [&]() -> int { return halo ? halo_allreduce_sum(halo, pcg_dot_slot(st), 1) : B2F_OK; };
   // q = A d and (d, q): one fused kernel when every integrator supports it, else two passes
   const bool fdot = operator_has_fused_dot(op);
   if (fdot && !op->dot_acc)
   {
      B2F_CUDA(cudaMalloc((void **)&op->dot_acc, sizeof(double) * 64));
      B2F_CUDA(cudaMemsetAsync(op->dot_acc, 0, sizeof(double) * 64, ctx->stream));
   }
   auto apply_and_dot = [&]() -> int
   {
      if (fdot)
      {
         B2F_TRY(mult_constrained(op, d, q, true, op->dot_acc));
         return collect_dot(ctx, op->dot_acc, pcg_dot_slot(st));
      }
      B2F_TRY(mult_constrained(op, d, q, true));
      return dot_to_device(ctx, n, d, q, pcg_dot_slot(st));
   };

   // r = b; x = 0   (iterative_mode = false, solvers.cpp:881-884)
   B2F_TRY(copy_vector(ctx, r, b, n));
   B2F_TRY(zero_vector(ctx, x, n));
   const double *zz = r; // "z" of the unpreconditioned iteration is r itself
   if (cheb)
   {
      B2F_TRY(chebyshev_apply(cheb, r, z));
      zz = z;
   }
   else if (dinv)
   {
      B2F_TRY(vector_mul(ctx, n, dinv, r, z)); // z = B r
      zz = z;
   }
   B2F_TRY(copy_vector(ctx, d, zz, n));
   B2F_TRY(dot_to_device(ctx, n, d, r, pcg_dot_slot(st)));
   B2F_TRY(reduce());
   B2F_TRY(pcg_after_nom0(ctx, st, opt->rel_tol, opt->abs_tol, opt->max_iter, hist,
                          hist ? history_capacity : 0));
   B2F_TRY(apply_and_dot()); // z = A d, den = (z, d)
   B2F_TRY(reduce());
   B2F_TRY(pcg_after_den0(ctx, st));

   unsigned char host_state[48 * sizeof(double)];
   int iter = 0, done = 0, converged = 0, final_iter = 0;
   double nom0 = 0, betanom = 0;
   auto poll = [&]() -> int
   {
      // (a store kernel into the pinned mirror, not a DMA copy: see publish_scalars)
      B2F_TRY(publish_scalars(ctx, (const double *)st, ctx->h_scalars, (pcg_state_bytes() + 7) / 8));
      memcpy(host_state, ctx->h_scalars, pcg_state_bytes());
      pcg_read_state(host_state, &iter, &done, &converged, &final_iter, &nom0, &betanom);
      return B2F_OK;
   };
   B2F_TRY(poll());
   int enq = 0; // iterations enqueued since the last poll
   while (!done)
   {
      B2F_TRY(pcg_update(ctx, st, n, x, d, r, q, dinv, z));
      if (cheb)
      {
         // general preconditioner: z = B r, betanom = (r, z) replace the fused Jacobi step
         B2F_TRY(chebyshev_apply(cheb, r, z));
         B2F_TRY(dot_to_device(ctx, n, r, z, pcg_dot_slot(st)));
      }
      B2F_TRY(reduce());
      B2F_TRY(pcg_after_betanom(ctx, st, hist, hist ? history_capacity : 0));
      B2F_TRY(pcg_direction(ctx, st, n, zz, d));
      B2F_TRY(apply_and_dot());
      B2F_TRY(reduce());
      B2F_TRY(pcg_after_den(ctx, st));
      if (++enq >= check_every) { B2F_TRY(poll()); enq = 0; }
   }
   result_h->iterations = final_iter;
   result_h->converged = converged;
   result_h->initial_norm = nom0 >= 0.0 ? sqrt(nom0) : nom0;
   result_h->final_norm = betanom >= 0.0 ? sqrt(betanom) : betanom;
   if (hist)
   {
      B2F_CUDA(cudaMemcpyAsync(history_h, hist, sizeof(double) * history_capacity,
                               cudaMemcpyDeviceToHost, ctx->stream));
      B2F_CUDA(cudaStreamSynchronize(ctx->stream));
   }
   return B2F_OK;
}

extern "C" {

int b2f_pcg(b2f_operator *op, const double *dinv, const double *b, double *x,
            const b2f_pcg_options *opt, b2f_pcg_result *result_h, double *history_h,
            int history_capacity)
{
   B2F_ON(op ? op->ctx : nullptr);
   return pcg_impl(op, dinv, nullptr, b, x, opt, result_h, history_h, history_capacity);
}

int b2f_pcg_chebyshev(b2f_operator *op, b2f_chebyshev *cheb, const double *b, double *x,
                      const b2f_pcg_options *opt, b2f_pcg_result *result_h, double *history_h,
                      int history_capacity)
{
   B2F_ON(op ? op->ctx : nullptr);
   B2F_REQUIRE(cheb && cheb->op == op, "b2f_pcg_chebyshev: smoother was built for another operator");
   return pcg_impl(op, nullptr, cheb, b, x, opt, result_h, history_h, history_capacity);
}

} // extern "C"
