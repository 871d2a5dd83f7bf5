// mfem_b200/csrc/restriction.cu -- ElementRestriction: L-vector <-> E-vector.
//
// Replaces ElementRestriction::{ctor, Mult, AbsMult, TAddMultTranspose, AbsMultTranspose}
// (fem/restriction.cpp:26-228).  These stand-alone kernels are bit-exact with the
// reference (same per-dof sequential summation order); the hot path does not use them --
// the fused operator kernel (pa_apply.cuh) gathers and scatters on the fly.
//
// HBM-bound integer/byte work: one thread per entry, coalesced index streams, random
// 8-byte gathers served from L2.
#include "common.cuh"
#include <cub/cub.cuh>

namespace b2f
{

template <bool ABS>
__global__ void __launch_bounds__(256)
gather_kernel(int64_t n, int nd, int vd, int byvdim, int ndofs, const int *__restrict__ gmap,
              const double *__restrict__ x, double *__restrict__ y)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   const int gid = __ldg(gmap + i);
   const bool plus = gid >= 0;
   const int j = plus ? gid : -1 - gid;
   const int64_t e = i / nd;
   const int d = (int)(i - e * nd);
   for (int c = 0; c < vd; ++c)
   {
      const double v = __ldg(x + (byvdim ? (int64_t)c + (int64_t)vd * j : (int64_t)j + (int64_t)ndofs * c));
      y[d + (int64_t)nd * (c + (int64_t)vd * e)] = (plus || ABS) ? v : -v;
   }
}

// MODE 0: set, 1: add, 2: abs-set
template <int MODE>
__global__ void __launch_bounds__(256)
scatter_kernel(int ndofs, int nd, int vd, int byvdim, const int *__restrict__ offsets,
               const int *__restrict__ indices, const double *__restrict__ x,
               double *__restrict__ y)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= ndofs) { return; }
   const int off = __ldg(offsets + i), next = __ldg(offsets + i + 1);
   for (int c = 0; c < vd; ++c)
   {
      double v = 0;
      for (int j = off; j < next; ++j)
      {
         const int s = __ldg(indices + j);
         const int idx = s >= 0 ? s : -1 - s;
         const int64_t e = idx / nd;
         const int d = idx - (int)e * nd;
         const double xv = __ldg(x + d + (int64_t)nd * (c + (int64_t)vd * e));
         v += (MODE == 2 || s >= 0) ? xv : -xv;
      }
      const int64_t o = byvdim ? (int64_t)c + (int64_t)vd * i : (int64_t)i + (int64_t)ndofs * c;
      if (MODE == 1) { y[o] += v; }
      else { y[o] = v; }
   }
}

// ---- device construction of the tables (restriction.cpp:71-113) ----------------------
__global__ void count_kernel(int64_t n, const int *__restrict__ elem_dof, int *offsets)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   const int s = elem_dof[i];
   atomicAdd(offsets + (s >= 0 ? s : -1 - s) + 1, 1);
}

__global__ void fill_kernel(int64_t n, int dof, const int *__restrict__ elem_dof,
                            const int *__restrict__ dof_map, const int *__restrict__ offsets,
                            int *cursor, int *gather_map, int *indices)
{
   const int64_t lid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (lid >= n) { return; }
   const int64_t e = lid / dof;
   const int d = (int)(lid - e * dof);
   const int sdid = dof_map ? dof_map[d] : 0;
   const int did = !dof_map ? d : (sdid >= 0 ? sdid : -1 - sdid);
   const int sgid = elem_dof[e * dof + did];
   const int gid = sgid >= 0 ? sgid : -1 - sgid;
   const bool plus = (sgid >= 0 && sdid >= 0) || (sgid < 0 && sdid < 0);
   gather_map[lid] = plus ? gid : -1 - gid;
   const int slot = atomicAdd(cursor + gid, 1);
   indices[offsets[gid] + slot] = plus ? (int)lid : -1 - (int)lid;
}

// the reference fills `indices` in ascending E-index order: sort each (short) segment
__global__ void sort_segments_kernel(int ndofs, const int *__restrict__ offsets, int *indices)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= ndofs) { return; }
   const int b = offsets[i], n = offsets[i + 1];
   for (int a = b + 1; a < n; ++a)
   {
      const int v = indices[a];
      const int kv = v >= 0 ? v : -1 - v;
      int j = a - 1;
      while (j >= b)
      {
         const int w = indices[j];
         if ((w >= 0 ? w : -1 - w) <= kv) { break; }
         indices[j + 1] = w;
         --j;
      }
      indices[j + 1] = v;
   }
}

int restriction_build_device(b2f_restriction *r, const int *elem_dof_d, const int *dof_map_d)
{
   b2f_ctx *ctx = r->ctx;
   const int64_t n = (int64_t)r->ne * r->dof;
   B2F_CUDA(cudaMemsetAsync(r->offsets, 0, sizeof(int) * (r->ndofs + 1), ctx->stream));
   if (n == 0) { return B2F_OK; }
   count_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(n, elem_dof_d, r->offsets);
   B2F_LAUNCH_CHECK(ctx);
   size_t tmp_bytes = 0;
   B2F_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, r->offsets + 1, r->offsets + 1,
                                          r->ndofs, ctx->stream));
   void *tmp = nullptr;
   int *cursor = nullptr;
   B2F_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
   B2F_CUDA(cudaMalloc(&cursor, sizeof(int) * (size_t)(r->ndofs + 1)));
   B2F_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)(r->ndofs + 1), ctx->stream));
   B2F_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, r->offsets + 1, r->offsets + 1,
                                          r->ndofs, ctx->stream));
   ctx->launches++;
   fill_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(n, r->dof, elem_dof_d, dof_map_d,
                                                                r->offsets, cursor,
                                                                r->gather_map, r->indices);
   B2F_LAUNCH_CHECK(ctx);
   sort_segments_kernel<<<(unsigned)cdiv(r->ndofs, 256), 256, 0, ctx->stream>>>(
      r->ndofs, r->offsets, r->indices);
   B2F_LAUNCH_CHECK(ctx);
   B2F_CUDA(cudaStreamSynchronize(ctx->stream));
   cudaFree(tmp);
   cudaFree(cursor);
   return B2F_OK;
}

int restriction_gather(b2f_restriction *r, const double *x, double *y, bool abs)
{
   b2f_ctx *ctx = r->ctx;
   const int64_t n = (int64_t)r->ne * r->dof;
   if (n == 0) { return B2F_OK; }
   const unsigned grid = (unsigned)cdiv(n, 256);
   if (abs)
   {
      gather_kernel<true><<<grid, 256, 0, ctx->stream>>>(n, r->dof, r->vdim, r->byvdim, r->ndofs,
                                                         r->gather_map, x, y);
   }
   else
   {
      gather_kernel<false><<<grid, 256, 0, ctx->stream>>>(n, r->dof, r->vdim, r->byvdim,
                                                          r->ndofs, r->gather_map, x, y);
   }
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

int restriction_scatter(b2f_restriction *r, const double *x, double *y, int mode)
{
   b2f_ctx *ctx = r->ctx;
   if (r->ndofs == 0) { return B2F_OK; }
   const unsigned grid = (unsigned)cdiv(r->ndofs, 256);
   switch (mode)
   {
      case 0:
         scatter_kernel<0><<<grid, 256, 0, ctx->stream>>>(r->ndofs, r->dof, r->vdim, r->byvdim,
                                                          r->offsets, r->indices, x, y);
         break;
      case 1:
         scatter_kernel<1><<<grid, 256, 0, ctx->stream>>>(r->ndofs, r->dof, r->vdim, r->byvdim,
                                                          r->offsets, r->indices, x, y);
         break;
      default:
         scatter_kernel<2><<<grid, 256, 0, ctx->stream>>>(r->ndofs, r->dof, r->vdim, r->byvdim,
                                                          r->offsets, r->indices, x, y);
   }
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

} // namespace b2f
