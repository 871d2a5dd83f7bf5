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
// Index rows: two 128-bit loads per thread at p = 1; at p = 2 (27 ints, not a whole number of words)
// the tile's rows are staged the same way -- read one int per lane at a 108-byte stride they cost 27
// LSU wavefronts per element, as many as everything else together (profiles/r02_ncu_bp1_p2_tpe.txt).
// With lane = element the gather and the reductions coalesce ACROSS elements (neighbours along x share
// consecutive vertex / edge / face / interior numbers): 5.9 + 3.4 wavefronts per element at p = 2.
// The z, D and z^T stages are fused per (qy,qx) column, in place, so the Q^3 array is never live.
// One thread per tile asks L2 for the tile the CTA `pf` blocks on will need (as pa_apply_eo.cuh).
#pragma once
#include "pa_eo_cfg.cuh"

namespace b2f
{

template <int D, int Q, int NT>
struct TpeCfg
{
   static constexpr int D3 = D * D * D, Q3 = Q * Q * Q;
   static constexpr int PITCH = Q3 | 1; // odd: lanes i, i+1 are PITCH doubles apart -> distinct banks
   // index rows that are not a whole number of 128-bit words (p = 2: 27 ints) would be read one int
   // per lane at a 4 D^3-byte stride -- 32 sectors per warp load, D^3 LSU wavefronts per element, as many
   // as the gather itself.  Those tiles are staged through shared memory instead (coalesced LDGSTS,
   // then conflict-free reads: the row pitch D^3 is odd).
   static constexpr bool STAGE_IDX = D3 % 4 != 0;
   static constexpr size_t SMEM = sizeof(double) * (size_t)NT * PITCH + (STAGE_IDX ? sizeof(int) * (size_t)NT * D3 : 0);
};

template <int D, int Q, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
pa_mass_tpe_kernel(const __grid_constant__ Basis<D, Q> bs, const double *__restrict__ qdata,
                   const double *__restrict__ x, double *__restrict__ y, const int *__restrict__ idx,
                   const int64_t ne, const int64_t e_first, const int pf)
{
   using C = TpeCfg<D, Q, NT>;
   constexpr int D3 = C::D3, Q3 = C::Q3, PITCH = C::PITCH;
   extern __shared__ __align__(16) double tsm[];
   const int t = threadIdx.x;
   const int64_t e0 = e_first + (int64_t)blockIdx.x * NT;
   const int nev = (int)((ne - e0) < NT ? (ne - e0) : NT);
   const bool act = t < nev;

   // gather indices first: they head the only dependent chain (index -> x)
   int *ism = reinterpret_cast<int *>(tsm + (size_t)NT * PITCH);
   if constexpr (C::STAGE_IDX)
   {
      const int *src = idx + e0 * (int64_t)D3;
      const int total = nev * D3;
      for (int i = t; i < total; i += NT) { cp_async4(ism + i, src + i); }
      cp_async_commit();
   }
   int gi[C::STAGE_IDX ? 1 : D3];
#pragma unroll
   for (int i = 0; i < (C::STAGE_IDX ? 1 : D3); i++) { gi[i] = -1; }
   if (act && !C::STAGE_IDX)
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
      cp_async_commit();
   }
   // ... and an L2 prefetch of the tile the CTA `pf` blocks on will ask for (the CTAs resident on the
   // device run roughly in block order, so that tile is wanted about one CTA lifetime from now)
   if (pf > 0 && t == 32 % NT)
   {
      const int64_t ep = e0 + (int64_t)pf * NT;
      if (ep + NT <= ne)
      {
         const uintptr_t q = reinterpret_cast<uintptr_t>(qdata + ep * (int64_t)Q3) & ~(uintptr_t)15;
         l2_prefetch_bulk(reinterpret_cast<const void *>(q), ((uint32_t)(NT * Q3) * 8u) & ~15u);
         const uintptr_t a = reinterpret_cast<uintptr_t>(idx + ep * D3) & ~(uintptr_t)15;
         l2_prefetch_bulk(reinterpret_cast<const void *>(a), ((uint32_t)(NT * D3) * 4u) & ~15u);
      }
   }
   double u[D][D][D]; // [dz][dy][dx]
   if constexpr (C::STAGE_IDX)
   {
      cp_async_wait_group<1>(); // the index tile; the quadrature data may still be in flight
      __syncthreads();
#pragma unroll
      for (int i = 0; i < D3; i++)
      {
         const int g = act ? ism[t * D3 + i] : -1;
         (&u[0][0][0])[i] = g >= 0 ? __ldg(x + g) : 0.0;
      }
   }
   else
   {
#pragma unroll
      for (int i = 0; i < D3; i++)
      {
         (&u[0][0][0])[i] = (act && gi[i] >= 0) ? __ldg(x + gi[i]) : 0.0;
      }
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
   // z, times the quadrature data, z^T: column by column, in place (a (qy,qx) column of b is D values
   // in, D values out, Q temporaries) -- the Q^3 array is never live, which is what bounds registers
#pragma unroll
   for (int qy = 0; qy < Q; qy++)
#pragma unroll
      for (int qx = 0; qx < Q; qx++)
      {
         double cz[Q];
#pragma unroll
         for (int qz = 0; qz < Q; qz++)
         {
            double s = 0.0;
#pragma unroll
            for (int dz = 0; dz < D; dz++) { s = fma(bs.B[qz][dz], b[dz][qy][qx], s); }
            cz[qz] = act ? s * dq[qx + Q * (qy + Q * qz)] : 0.0;
         }
#pragma unroll
         for (int dz = 0; dz < D; dz++)
         {
            double s = 0.0;
#pragma unroll
            for (int qz = 0; qz < Q; qz++) { s = fma(bs.B[qz][dz], cz[qz], s); }
            b[dz][qy][qx] = s;
         }
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
            int g;
            if constexpr (C::STAGE_IDX) { g = act ? ism[t * D3 + dx + D * (dy + D * dz)] : -1; }
            else { g = gi[dx + D * (dy + D * dz)]; }
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
   const char *pfs = getenv("B2F_PF"); // prefetch distance in tiles; default: the CTAs resident on the device
   const int pf = pfs ? atoi(pfs) : resident;
   timing_mark(ctx);
   kern<<<(unsigned)grid, NT, C::SMEM, ctx->stream>>>(bs, pa->data, x, y, idx, e_last, e_first, pf);
   timing_mark(ctx);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

} // namespace b2f
