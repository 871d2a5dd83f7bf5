// mfem_b200/csrc/pa_apply_tpe.cuh -- the lowest orders: one THREAD per element, no shared-memory
// contraction stages at all.
//
// Same operator as pa_apply_eo.cuh (MFEM's SmemPAMassApply3D, fem/integ/bilininteg_mass_kernels.hpp:
// 815-1165, fused with ElementRestriction::Mult / MultTranspose, fem/restriction.cpp:116,159).
// Why a second formulation: the five-stage pencil kernel moves every intermediate through shared
// memory (4 Q D^2 + 4 Q^2 D + Q^3 doubles per element), and for the mass operator that traffic plus
// the gather and the reductions saturates the LSU data pipe long before HBM (DESIGN.md 4.1:
// l1tex__data_pipe_lsu_wavefronts 96 %).  At p = 1, 2 an element is 8 / 27 dofs and 27 / 64
// quadrature points: one thread holds all of it in registers, the three contractions, the
// pointwise multiply and the three transposed contractions are straight-line FMAs with the basis as
// constant-bank operands, and the only LSU work left is what the operator must do anyway -- index
// row, gather, quadrature data, scatter.  No barrier except the one that publishes the tile's
// quadrature data.
//
// Quadrature data: the tile's NT element slabs are one contiguous block in HBM; they are copied
// cooperatively with cp.async (coalesced 8-byte LDGSTS) into shared memory with an ODD per-element
// pitch, so that the per-thread reads (lane i reads point q of element i) are conflict-free.
#pragma once
#include "pa_eo_cfg.cuh"

namespace b2f
{

template <int D, int Q, int NT>
struct TpeCfg
{
   static constexpr int D3 = D * D * D, Q3 = Q * Q * Q;
   static constexpr int PITCH = Q3 | 1; // odd: lanes i, i+1 are PITCH doubles apart -> distinct banks
   static constexpr size_t SMEM = sizeof(double) * (size_t)NT * PITCH;
};

template <int D, int Q, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
pa_mass_tpe_kernel(const __grid_constant__ Basis<D, Q> bs, const double *__restrict__ qdata,
                   const double *__restrict__ x, double *__restrict__ y, const int *__restrict__ idx,
                   const int64_t ne, const int64_t e_first)
{
   using C = TpeCfg<D, Q, NT>;
   constexpr int D3 = C::D3, Q3 = C::Q3, PITCH = C::PITCH;
   extern __shared__ __align__(16) double tsm[];
   const int t = threadIdx.x;
   const int64_t e0 = e_first + (int64_t)blockIdx.x * NT;
   const int nev = (int)((ne - e0) < NT ? (ne - e0) : NT);
   const bool act = t < nev;

   // gather indices first: they head the only dependent chain (index -> x)
   int gi[D3];
#pragma unroll
   for (int i = 0; i < D3; i++) { gi[i] = -1; }
   if (act)
   {
      const int *ip = idx + (e0 + t) * D3;
      if constexpr (D3 % 4 == 0)
      {
         const int4 *ip4 = reinterpret_cast<const int4 *>(ip); // a row of D3 ints is 16-byte aligned
#pragma unroll
         for (int i = 0; i < D3 / 4; i++)
         {
            const int4 v = __ldg(ip4 + i);
            gi[4 * i] = v.x; gi[4 * i + 1] = v.y; gi[4 * i + 2] = v.z; gi[4 * i + 3] = v.w;
         }
      }
      else
      {
#pragma unroll
         for (int i = 0; i < D3; i++) { gi[i] = __ldg(ip + i); }
      }
   }
   // the tile's quadrature data: coalesced asynchronous copy into odd-pitch slots
   {
      const double *src = qdata + e0 * (int64_t)Q3;
      const int total = nev * Q3;
      for (int i = t; i < total; i += NT)
      {
         const int el = i / Q3;
         cp_async8(tsm + el * PITCH + (i - el * Q3), src + i);
      }
   }
   double u[D][D][D]; // [dz][dy][dx]
#pragma unroll
   for (int i = 0; i < D3; i++)
   {
      (&u[0][0][0])[i] = (act && gi[i] >= 0) ? __ldg(x + gi[i]) : 0.0;
   }
   // x: D -> Q
   double a[D][D][Q];
#pragma unroll
   for (int dz = 0; dz < D; dz++)
#pragma unroll
      for (int dy = 0; dy < D; dy++)
#pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            double s = 0.0;
#pragma unroll
            for (int dx = 0; dx < D; dx++) { s = fma(bs.B[qx][dx], u[dz][dy][dx], s); }
            a[dz][dy][qx] = s;
         }
   // y
   double b[D][Q][Q];
#pragma unroll
   for (int dz = 0; dz < D; dz++)
#pragma unroll
      for (int qy = 0; qy < Q; qy++)
#pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            double s = 0.0;
#pragma unroll
            for (int dy = 0; dy < D; dy++) { s = fma(bs.B[qy][dy], a[dz][dy][qx], s); }
            b[dz][qy][qx] = s;
         }
   // z, times the quadrature data
   cp_async_wait_all();
   __syncthreads();
   const double *dq = tsm + t * PITCH;
   double c[Q][Q][Q];
#pragma unroll
   for (int qz = 0; qz < Q; qz++)
#pragma unroll
      for (int qy = 0; qy < Q; qy++)
#pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            double s = 0.0;
#pragma unroll
            for (int dz = 0; dz < D; dz++) { s = fma(bs.B[qz][dz], b[dz][qy][qx], s); }
            c[qz][qy][qx] = act ? s * dq[qx + Q * (qy + Q * qz)] : 0.0;
         }
   // z^T
#pragma unroll
   for (int dz = 0; dz < D; dz++)
#pragma unroll
      for (int qy = 0; qy < Q; qy++)
#pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            double s = 0.0;
#pragma unroll
            for (int qz = 0; qz < Q; qz++) { s = fma(bs.B[qz][dz], c[qz][qy][qx], s); }
            b[dz][qy][qx] = s;
         }
   // y^T
#pragma unroll
   for (int dz = 0; dz < D; dz++)
#pragma unroll
      for (int dy = 0; dy < D; dy++)
#pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            double s = 0.0;
#pragma unroll
            for (int qy = 0; qy < Q; qy++) { s = fma(bs.B[qy][dy], b[dz][qy][qx], s); }
            a[dz][dy][qx] = s;
         }
   // x^T and the scatter
#pragma unroll
   for (int dz = 0; dz < D; dz++)
#pragma unroll
      for (int dy = 0; dy < D; dy++)
#pragma unroll
         for (int dx = 0; dx < D; dx++)
         {
            double s = 0.0;
#pragma unroll
            for (int qx = 0; qx < Q; qx++) { s = fma(bs.B[qx][dx], a[dz][dy][qx], s); }
            const int g = gi[dx + D * (dy + D * dz)];
            if (g >= 0) { atomicAdd(y + g, s); }
         }
}

// fused L-vector form only (the E-vector seam keeps the pencil kernel); element range as launch_3d_eo
template <int D, int Q, int NT = 128, int MINB = 2>
int launch_mass_tpe(b2f_pa *pa, const double *x, double *y, const int *idx)
{
   using C = TpeCfg<D, Q, NT>;
   b2f_ctx *ctx = pa->ctx;
   Basis<D, Q> bs;
   fill_basis(pa, bs);
   const int64_t e_first = pa->rb, e_last = pa->re < 0 ? pa->ne : pa->re;
   if (e_last <= e_first) { return B2F_OK; }
   const int64_t grid = cdiv(e_last - e_first, NT);
   auto kern = pa_mass_tpe_kernel<D, Q, NT, MINB>;
   int resident = 0;
   if (int rc = kernel_residency(ctx, kern, NT, C::SMEM, resident)) { return rc; }
   timing_mark(ctx);
   kern<<<(unsigned)grid, NT, C::SMEM, ctx->stream>>>(bs, pa->data, x, y, idx, e_last, e_first);
   timing_mark(ctx);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

} // namespace b2f
