// mfem_b200/csrc/vector_ops.cu -- Vector BLAS-1 and the fused PCG vector kernels.
//
// Replaces Vector::operator= / add / subtract / operator* (linalg/vector.cpp:436,570,1079),
// the block-tree reduction that stages partials through pinned host memory and syncs the
// device on every dot (general/reducers.hpp:451-594), OperatorJacobiSmoother::Mult's three
// passes (linalg/solvers.cpp:427-450) and the scalar bookkeeping of CGSolver::Mult
// (linalg/solvers.cpp:869-1050).
//
// B200 design: HBM-bound streaming kernels, 2 x 128-bit loads in flight per thread,
// warp-shuffle reductions, grid = a multiple of the SM count; every PCG scalar
// (nom, den, betanom, alpha, beta, iteration state) lives in device memory, so an
// iteration enqueues without a host round trip.
#include "common.cuh"

namespace b2f
{

constexpr int VT = 256; // threads per block for streaming kernels

static inline unsigned stream_grid(const b2f_ctx *ctx, int64_t n, int per_thread)
{
   int64_t want = cdiv(n, (int64_t)VT * per_thread);
   int64_t cap = (int64_t)ctx->sm_count * 8;
   if (want > cap) { want = cap; }
   if (want < 1) { want = 1; }
   return (unsigned)want;
}

__global__ void __launch_bounds__(VT) fill_kernel(double *x, double v, int64_t n)
{
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      x[i] = v;
   }
}

__global__ void set_entries_kernel(double *x, const int *__restrict__ idx, int n, double v)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { x[idx[i]] = v; }
}

__global__ void copy_entries_kernel(double *dst, const double *__restrict__ src,
                                    const int *__restrict__ idx, int n)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { const int k = idx[i]; dst[k] = src[k]; }
}

__global__ void __launch_bounds__(256)
copy_entries_sqsum_kernel(double *dst, const double *__restrict__ src, const int *__restrict__ idx,
                          int n, double *acc)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   double v = 0.0;
   if (i < n) { const int k = idx[i]; v = src[k]; dst[k] = v; }
   v *= v;
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
   if ((threadIdx.x & 31) == 0 && v != 0.0) { atomicAdd(acc + (blockIdx.x & 63), v); }
}

__global__ void collect_dot_kernel(double *acc, double *out)
{
   double v = acc[threadIdx.x] + acc[threadIdx.x + 32];
   acc[threadIdx.x] = 0.0;
   acc[threadIdx.x + 32] = 0.0;
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
   if (threadIdx.x == 0) { *out = v; }
}

__global__ void __launch_bounds__(VT)
add_kernel(int64_t n, const double *__restrict__ x, double a, const double *__restrict__ y,
           double *z)
{
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      z[i] = x[i] + a * y[i];
   }
}

__global__ void __launch_bounds__(VT)
mul_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ x, double *y)
{
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      y[i] = 0.0 + a[i] * x[i]; // y = 0; y += dinv*r  (solvers.cpp:441-449)
   }
}

// One term of OperatorChebyshevSmoother::Mult (linalg/solvers.cpp:637-657):
// R = dinv .* src; y (+)= c R.   FIRST: y starts from 0 (the reference's y = 0.0).
template <bool FIRST>
__global__ void __launch_bounds__(VT)
cheb_step_kernel(int64_t n, const double *__restrict__ src, const double *__restrict__ dinv,
                 double c, double *__restrict__ res, double *__restrict__ y)
{
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      const double r = src[i] * dinv[i];
      res[i] = r;
      y[i] = (FIRST ? 0.0 : y[i]) + c * r;
   }
}

// v *= s, s = 1 / sqrt(*d_normsq)   (PowerMethod, linalg/operator.cpp:913)
__global__ void __launch_bounds__(VT)
normalize_kernel(int64_t n, const double *__restrict__ d_normsq, double *v)
{
   const double s = sqrt(*d_normsq);
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      v[i] /= s;
   }
}

__global__ void jacobi_setup_kernel(int n, const double *__restrict__ diag, double damping,
                                    double *dinv)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { dinv[i] = damping / diag[i]; }
}

// ---- block reduction: warp shuffles, then one shuffle pass over the warp sums --------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
   return v;
}

// Sum `v` over the grid into *out.  Block partials go to `partials`; the last block to
// finish (ticket counter) adds them in index order, so the result is deterministic for a
// fixed grid size.
__device__ __forceinline__ void grid_sum(double v, double *partials, unsigned int *counter,
                                         double *out)
{
   __shared__ double wsum[VT / 32];
   __shared__ bool last;
   v = warp_sum(v);
   if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = v; }
   __syncthreads();
   if (threadIdx.x < 32)
   {
      double s = threadIdx.x < VT / 32 ? wsum[threadIdx.x] : 0.0;
      s = warp_sum(s);
      if (threadIdx.x == 0)
      {
         partials[blockIdx.x] = s;
         __threadfence();
         const unsigned int ticket = atomicInc(counter, gridDim.x - 1);
         last = (ticket == gridDim.x - 1);
      }
   }
   __syncthreads();
   if (last && threadIdx.x < 32)
   {
      __threadfence();
      double s = 0.0;
      for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) { s += partials[i]; }
      s = warp_sum(s);
      if (threadIdx.x == 0) { *out = s; }
   }
}

__global__ void __launch_bounds__(VT)
dot_kernel(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
           double *partials, unsigned int *counter, double *out)
{
   double s = 0.0;
   const int64_t stride = (int64_t)gridDim.x * VT;
   int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x;
   for (; i + stride < n; i += 2 * stride)
   {
      const double a0 = x[i], b0 = y[i], a1 = x[i + stride], b1 = y[i + stride];
      s = fma(a0, b0, s);
      s = fma(a1, b1, s);
   }
   if (i < n) { s = fma(x[i], y[i], s); }
   grid_sum(s, partials, counter, out);
}

// ---- PCG (linalg/solvers.cpp:869-1050) ------------------------------------------------
// device-resident solver state; one instance lives in ctx->d_scalars
struct PcgState
{
   double nom, den, betanom, beta, r0, nom0;
   double dot_tmp;
   int iter;       // the reference's loop counter i
   int done;       // 1: loop exited
   int converged;
   int final_iter;
   int max_iter;
};

// x += alpha d ; r -= alpha q ; z = dinv r (or z aliases r) ; partial (r, z)
__global__ void __launch_bounds__(VT)
pcg_update_kernel(int64_t n, PcgState *st, double *x, const double *__restrict__ d, double *r,
                  const double *__restrict__ q, const double *__restrict__ dinv, double *z,
                  double *partials, unsigned int *counter)
{
   if (st->done) { return; }
   const double alpha = st->nom / st->den;
   double s = 0.0;
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      x[i] = x[i] + alpha * d[i];
      const double ri = r[i] + (-alpha) * q[i];
      r[i] = ri;
      if (dinv)
      {
         const double zi = 0.0 + dinv[i] * ri;
         z[i] = zi;
         s = fma(ri, zi, s);
      }
      else { s = fma(ri, ri, s); }
   }
   grid_sum(s, partials, counter, &st->dot_tmp);
}

// 128-bit form of the same update (all vectors 16-byte aligned): two entries per access
template <bool PREC>
__global__ void __launch_bounds__(VT)
pcg_update2_kernel(int64_t n, PcgState *st, double *x, const double *__restrict__ d, double *r,
                   const double *__restrict__ q, const double *__restrict__ dinv, double *z,
                   double *partials, unsigned int *counter)
{
   if (st->done) { return; }
   const double alpha = st->nom / st->den;
   double s = 0.0;
   const int64_t n2 = n >> 1;
   double2 *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r);
   double2 *z2 = reinterpret_cast<double2 *>(z);
   const double2 *d2 = reinterpret_cast<const double2 *>(d), *q2 = reinterpret_cast<const double2 *>(q);
   const double2 *p2 = reinterpret_cast<const double2 *>(dinv);
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * VT)
   {
      double2 xv = x2[i], rv = r2[i];
      const double2 dv = d2[i], qv = q2[i];
      xv.x = xv.x + alpha * dv.x; xv.y = xv.y + alpha * dv.y;
      rv.x = rv.x + (-alpha) * qv.x; rv.y = rv.y + (-alpha) * qv.y;
      x2[i] = xv;
      r2[i] = rv;
      if (PREC)
      {
         const double2 pv = p2[i];
         double2 zv;
         zv.x = 0.0 + pv.x * rv.x; zv.y = 0.0 + pv.y * rv.y;
         z2[i] = zv;
         s = fma(rv.x, zv.x, s);
         s = fma(rv.y, zv.y, s);
      }
      else
      {
         s = fma(rv.x, rv.x, s);
         s = fma(rv.y, rv.y, s);
      }
   }
   if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
   {
      const int64_t i = n - 1;
      x[i] = x[i] + alpha * d[i];
      const double ri = r[i] + (-alpha) * q[i];
      r[i] = ri;
      if (PREC) { const double zi = 0.0 + dinv[i] * ri; z[i] = zi; s = fma(ri, zi, s); }
      else { s = fma(ri, ri, s); }
   }
   grid_sum(s, partials, counter, &st->dot_tmp);
}

__global__ void __launch_bounds__(VT)
pcg_direction2_kernel(int64_t n, const PcgState *st, const double *__restrict__ z, double *d)
{
   if (st->done) { return; }
   const double beta = st->beta;
   const int64_t n2 = n >> 1;
   const double2 *z2 = reinterpret_cast<const double2 *>(z);
   double2 *d2 = reinterpret_cast<double2 *>(d);
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * VT)
   {
      const double2 zv = z2[i];
      double2 dv = d2[i];
      dv.x = zv.x + beta * dv.x; dv.y = zv.y + beta * dv.y;
      d2[i] = dv;
   }
   if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) { d[n - 1] = z[n - 1] + beta * d[n - 1]; }
}

__global__ void __launch_bounds__(VT)
dot2_kernel(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
            double *partials, unsigned int *counter, double *out)
{
   double s0 = 0.0, s1 = 0.0;
   const int64_t n2 = n >> 1;
   const double2 *x2 = reinterpret_cast<const double2 *>(x), *y2 = reinterpret_cast<const double2 *>(y);
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * VT)
   {
      const double2 a = x2[i], b = y2[i];
      s0 = fma(a.x, b.x, s0);
      s1 = fma(a.y, b.y, s1);
   }
   if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) { s0 = fma(x[n - 1], y[n - 1], s0); }
   grid_sum(s0 + s1, partials, counter, out);
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// scalar control flow after betanom = (r, z) is known (solvers.cpp:963-1000)
__global__ void pcg_after_betanom_kernel(PcgState *st, double *history, int history_cap)
{
   if (st->done) { return; }
   const double betanom = st->dot_tmp;
   st->betanom = betanom;
   const int i = st->iter;
   if (history && i < history_cap) { history[i] = betanom; }
   if (betanom < 0.0) { st->done = 1; st->converged = 0; st->final_iter = i; return; }
   if (betanom <= st->r0) { st->done = 1; st->converged = 1; st->final_iter = i; return; }
   st->iter = i + 1;
   if (i + 1 > st->max_iter) { st->done = 1; st->converged = 0; st->final_iter = st->max_iter; return; }
   st->beta = betanom / st->nom;
}

// d = z + beta d   (z = r when unpreconditioned)
__global__ void __launch_bounds__(VT)
pcg_direction_kernel(int64_t n, const PcgState *st, const double *__restrict__ z, double *d)
{
   if (st->done) { return; }
   const double beta = st->beta;
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT)
   {
      d[i] = z[i] + beta * d[i];
   }
}

// den = (d, q) is known (solvers.cpp:1010-1024)
__global__ void pcg_after_den_kernel(PcgState *st)
{
   if (st->done) { return; }
   const double den = st->dot_tmp;
   st->den = den;
   if (den == 0.0) { st->done = 1; st->final_iter = st->iter; return; }
   st->nom = st->betanom;
}

// start-up: nom0 = nom = (d, r) known (solvers.cpp:894-935)
__global__ void pcg_after_nom0_kernel(PcgState *st, double rel_tol, double abs_tol, int max_iter,
                                      double *history, int history_cap)
{
   const double nom = st->dot_tmp;
   st->nom = st->nom0 = st->betanom = nom;
   st->iter = 1; st->done = 0; st->converged = 0; st->final_iter = max_iter;
   st->max_iter = max_iter;
   if (history && history_cap > 0) { history[0] = nom; }
   if (nom < 0.0) { st->done = 1; st->final_iter = 0; return; }
   const double a = nom * rel_tol * rel_tol, b = abs_tol * abs_tol;
   st->r0 = a > b ? a : b;
   if (nom <= st->r0) { st->done = 1; st->converged = 1; st->final_iter = 0; }
}

__global__ void pcg_after_den0_kernel(PcgState *st)
{
   if (st->done) { return; }
   st->den = st->dot_tmp;
   if (st->den == 0.0) { st->done = 1; st->converged = 0; st->final_iter = 0; }
}

// zero / copy as KERNELS (128-bit accesses).  cudaMemsetAsync / cudaMemcpyAsync may be served by a
// copy engine, where they queue behind whatever host<->device transfers the application has in
// flight (bench.py's double-buffered end-to-end path): the solver's per-iteration zero fill and its
// start-up copies must not wait for a 0.8 GB download.
__global__ void __launch_bounds__(VT) zero_kernel(double *__restrict__ x, int64_t n)
{
   const int64_t n2 = n / 2;
   double2 *x2 = reinterpret_cast<double2 *>(x);
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * VT)
   {
      x2[i] = make_double2(0.0, 0.0);
   }
   if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) { x[n - 1] = 0.0; }
}

__global__ void __launch_bounds__(VT) copy_kernel(double *__restrict__ dst, const double *__restrict__ src, int64_t n)
{
   const int64_t n2 = n / 2;
   double2 *d2 = reinterpret_cast<double2 *>(dst);
   const double2 *s2 = reinterpret_cast<const double2 *>(src);
   for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * VT) { d2[i] = s2[i]; }
   if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) { dst[n - 1] = src[n - 1]; }
}

// ---- host launchers --------------------------------------------------------------------
int zero_vector(b2f_ctx *ctx, double *x, int64_t n)
{
   if (n == 0) { return B2F_OK; }
   if (reinterpret_cast<uintptr_t>(x) & 15) { return fill(ctx, x, 0.0, n); } // (never for cudaMalloc'ed vectors)
   zero_kernel<<<stream_grid(ctx, n / 2 + 1, 4), VT, 0, ctx->stream>>>(x, n);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int copy_vector(b2f_ctx *ctx, double *dst, const double *src, int64_t n)
{
   if (n == 0 || dst == src) { return B2F_OK; }
   if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15)
   {
      B2F_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
      return B2F_OK;
   }
   copy_kernel<<<stream_grid(ctx, n / 2 + 1, 4), VT, 0, ctx->stream>>>(dst, src, n);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int fill(b2f_ctx *ctx, double *x, double v, int64_t n)
{
   if (n == 0) { return B2F_OK; }
   if (v == 0.0)
   {
      B2F_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * n, ctx->stream));
      return B2F_OK;
   }
   fill_kernel<<<stream_grid(ctx, n, 4), VT, 0, ctx->stream>>>(x, v, n);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int set_entries(b2f_ctx *ctx, double *x, const int *idx, int n, double v)
{
   if (n == 0) { return B2F_OK; }
   set_entries_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(x, idx, n, v);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int copy_entries(b2f_ctx *ctx, double *dst, const double *src, const int *idx, int n)
{
   if (n == 0) { return B2F_OK; }
   copy_entries_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(dst, src, idx, n);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int copy_entries_sqsum(b2f_ctx *ctx, double *dst, const double *src, const int *idx, int n, double *acc)
{
   if (n == 0) { return B2F_OK; }
   copy_entries_sqsum_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(dst, src, idx, n, acc);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int collect_dot(b2f_ctx *ctx, double *acc, double *out)
{
   collect_dot_kernel<<<1, 32, 0, ctx->stream>>>(acc, out);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int dot_to_device(b2f_ctx *ctx, int64_t n, const double *x, const double *y, double *d_out)
{
   unsigned grid = stream_grid(ctx, n, 8);
   if ((int)grid > ctx->n_partials) { grid = ctx->n_partials; }
   if (aligned16(x) && aligned16(y) && n >= 2)
   {
      dot2_kernel<<<grid, VT, 0, ctx->stream>>>(n, x, y, ctx->d_partials, ctx->d_counter, d_out);
      B2F_LAUNCH_CHECK(ctx);
      return B2F_OK;
   }
   dot_kernel<<<grid, VT, 0, ctx->stream>>>(n, x, y, ctx->d_partials, ctx->d_counter, d_out);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int vector_add(b2f_ctx *ctx, int64_t n, const double *x, double a, const double *y, double *z)
{
   if (n == 0) { return B2F_OK; }
   add_kernel<<<stream_grid(ctx, n, 4), VT, 0, ctx->stream>>>(n, x, a, y, z);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int vector_mul(b2f_ctx *ctx, int64_t n, const double *a, const double *x, double *y)
{
   if (n == 0) { return B2F_OK; }
   mul_kernel<<<stream_grid(ctx, n, 4), VT, 0, ctx->stream>>>(n, a, x, y);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int jacobi_setup(b2f_ctx *ctx, int n, const double *diag, const int *ess, int n_ess,
                 double damping, double *dinv)
{
   if (n == 0) { return B2F_OK; }
   jacobi_setup_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(n, diag, damping, dinv);
   B2F_LAUNCH_CHECK(ctx);
   return set_entries(ctx, dinv, ess, n_ess, damping);
}

int cheb_step(b2f_ctx *ctx, int64_t n, const double *src, const double *dinv, double c, double *res,
              double *y, bool first)
{
   if (n == 0) { return B2F_OK; }
   if (first) { cheb_step_kernel<true><<<stream_grid(ctx, n, 4), VT, 0, ctx->stream>>>(n, src, dinv, c, res, y); }
   else { cheb_step_kernel<false><<<stream_grid(ctx, n, 4), VT, 0, ctx->stream>>>(n, src, dinv, c, res, y); }
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int normalize_by(b2f_ctx *ctx, int64_t n, const double *d_normsq, double *v)
{
   if (n == 0) { return B2F_OK; }
   normalize_kernel<<<stream_grid(ctx, n, 4), VT, 0, ctx->stream>>>(n, d_normsq, v);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

// pieces of the PCG loop, called from capi.cu
int pcg_state_bytes() { return (int)sizeof(PcgState); }

// Device scalars -> the context's pinned host mirror WITHOUT a DMA copy: a one-warp kernel stores
// straight into the (UVA-mapped) pinned buffer.  A cudaMemcpyAsync of a few bytes would queue on
// the device-to-host copy engine behind whatever the application has in flight there -- with a
// 0.8 GB solution download of the previous solve running, every host poll of the PCG state stalled
// the solver for the rest of that download (the end-to-end numbers of bench.py at 4 and 8 GPUs).
__global__ void publish_scalars_kernel(const double *__restrict__ src, volatile double *dst_host, int n)
{
   for (int i = threadIdx.x; i < n; i += blockDim.x) { dst_host[i] = src[i]; }
   __threadfence_system();
}

int publish_scalars(b2f_ctx *ctx, const double *d_src, double *h_dst, int n)
{
   publish_scalars_kernel<<<1, 32, 0, ctx->stream>>>(d_src, h_dst, n);
   B2F_LAUNCH_CHECK(ctx);
   B2F_CUDA(cudaStreamSynchronize(ctx->stream));
   return B2F_OK;
}

int pcg_start(b2f_ctx *ctx, void *state, int64_t n, const double *d, const double *r,
              double rel_tol, double abs_tol, int max_iter, double *history, int history_cap,
              double *dot_out_for_allreduce)
{
   PcgState *st = (PcgState *)state;
   (void)dot_out_for_allreduce;
   int rc = dot_to_device(ctx, n, d, r, &st->dot_tmp);
   if (rc) { return rc; }
   return B2F_OK;
}

int pcg_after_nom0(b2f_ctx *ctx, void *state, double rel_tol, double abs_tol, int max_iter,
                   double *history, int history_cap)
{
   pcg_after_nom0_kernel<<<1, 1, 0, ctx->stream>>>((PcgState *)state, rel_tol, abs_tol, max_iter,
                                                   history, history_cap);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int pcg_after_den0(b2f_ctx *ctx, void *state)
{
   pcg_after_den0_kernel<<<1, 1, 0, ctx->stream>>>((PcgState *)state);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int pcg_update(b2f_ctx *ctx, void *state, int64_t n, double *x, const double *d, double *r,
               const double *q, const double *dinv, double *z)
{
   unsigned grid = stream_grid(ctx, n, 4);
   if ((int)grid > ctx->n_partials) { grid = ctx->n_partials; }
   if (aligned16(x) && aligned16(d) && aligned16(r) && aligned16(q) && aligned16(dinv) && aligned16(z) && n >= 2)
   {
      if (dinv)
      {
         pcg_update2_kernel<true><<<grid, VT, 0, ctx->stream>>>(n, (PcgState *)state, x, d, r, q, dinv, z,
                                                                ctx->d_partials, ctx->d_counter);
      }
      else
      {
         pcg_update2_kernel<false><<<grid, VT, 0, ctx->stream>>>(n, (PcgState *)state, x, d, r, q, dinv, z,
                                                                 ctx->d_partials, ctx->d_counter);
      }
      B2F_LAUNCH_CHECK(ctx);
      return B2F_OK;
   }
   pcg_update_kernel<<<grid, VT, 0, ctx->stream>>>(n, (PcgState *)state, x, d, r, q, dinv, z,
                                                   ctx->d_partials, ctx->d_counter);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int pcg_after_betanom(b2f_ctx *ctx, void *state, double *history, int history_cap)
{
   pcg_after_betanom_kernel<<<1, 1, 0, ctx->stream>>>((PcgState *)state, history, history_cap);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int pcg_direction(b2f_ctx *ctx, void *state, int64_t n, const double *z, double *d)
{
   if (aligned16(z) && aligned16(d) && n >= 2)
   {
      pcg_direction2_kernel<<<stream_grid(ctx, n, 4), VT, 0, ctx->stream>>>(n, (PcgState *)state, z, d);
      B2F_LAUNCH_CHECK(ctx);
      return B2F_OK;
   }
   pcg_direction_kernel<<<stream_grid(ctx, n, 4), VT, 0, ctx->stream>>>(n, (PcgState *)state, z, d);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int pcg_after_den(b2f_ctx *ctx, void *state)
{
   pcg_after_den_kernel<<<1, 1, 0, ctx->stream>>>((PcgState *)state);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

double *pcg_dot_slot(void *state) { return &((PcgState *)state)->dot_tmp; }

void pcg_read_state(const void *host_copy, int *iter, int *done, int *converged, int *final_iter,
                    double *nom0, double *betanom)
{
   const PcgState *st = (const PcgState *)host_copy;
   *iter = st->iter; *done = st->done; *converged = st->converged; *final_iter = st->final_iter;
   *nom0 = st->nom0; *betanom = st->betanom;
}

} // namespace b2f
