// mfem_b200/csrc/pa_apply_eo_lean.cuh -- EXPERIMENTAL (B2F_VARIANT=60, tools/opbench.py --variants v60):
// the even-odd fused kernel of pa_apply_eo.cuh with a shorter prologue.
//
// profiles/r01_pa_apply_3d_eo_p4_n133_final_full.txt: 250 of the 1033 warp instructions issued per
// p=4 diffusion element (212 of 567 for mass) run before the first barrier, and only 46 (25) of
// them are FP64.  What this variant changes, nothing else (same stage functions, same layout, same
// arithmetic, bit-identical results):
//   * stage-1/5 lane coordinates (element, offset inside the element, A/F address) come packed in
//     one word from a kernel-parameter table instead of div/mod chains by NEB and D;
//   * warps without a stage-1 line skip the whole gather/stage-1 block;
//   * the TMA issue and the L2 prefetch run in such a warp when there is one (they were on the
//     critical path of warps 0 and 1), and the full-tile case is unrolled at compile time.
// Fused scalar form only (no vector, no fused dot, no collocated shortcut).  Not the default:
// it has not been validated on hardware yet.
#pragma once
#include "pa_apply_eo.cuh"

namespace b2f
{

template <int N>
struct LaneTab
{
   unsigned v[N]; // a1 | loc1 << 16 | e1 << 26
};

template <int KIND, int D, int Q, int NEB, int MINB>
__global__ void __launch_bounds__(PipeCfg<KIND, D, Q, NEB>::NT, MINB)
pa_apply_3d_eo_lean_kernel(const __grid_constant__ EOBasis<D, Q> bs,
                           const __grid_constant__ LaneTab<((PipeCfg<KIND, D, Q, NEB>::T1 + 3) / 4) * 4> lt,
                           const double *__restrict__ qdata, const double *__restrict__ x,
                           double *__restrict__ y, const int *__restrict__ idx, const int64_t ne,
                           const int pf, const int64_t e_first)
{
   using C = PipeCfg<KIND, D, Q, NEB>;
   using S = EOStages<KIND, D, Q, NEB, false, false>;
   static_assert(C::TMA_OK, "the lean variant streams the quadrature data with TMA");
   constexpr int D3 = C::D3, Q3 = C::Q3;
   // a warp without stage-1 lines issues the copies; warp 0 when every warp has lines
   constexpr int ISSUER = (C::NT - 32 >= C::T1) ? C::NT - 32 : 0;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
   double *qd = reinterpret_cast<double *>(smem_raw + 16);
   double *r1 = qd + C::QD;
   double *r2 = r1 + C::R1;

   const int t = threadIdx.x;
   const int64_t e0 = e_first + (int64_t)blockIdx.x * NEB; // elements [e_first, ne) of this launch
   const int nev = (int)((ne - e0) < NEB ? (ne - e0) : NEB);

   const bool has1 = t < C::T1;
   bool act1 = false;
   unsigned lc = 0;
   int gi[D];
#pragma unroll
   for (int dy = 0; dy < D; dy++) { gi[dy] = -1; }
   if (has1)
   {
      lc = lt.v[t];
      act1 = (int)(lc >> 26) < nev;
      if (act1)
      {
         const int *ip = idx + (e0 + (lc >> 26)) * D3 + ((lc >> 16) & 1023u);
#pragma unroll
         for (int dy = 0; dy < D; dy++) { gi[dy] = __ldg(ip + D * dy); }
      }
   }
   if (t == ISSUER)
   {
      const double *qslab = qdata + e0 * (int64_t)(C::NQD * Q3);
      mbar_init(bar, 1);
      if (nev == NEB)
      {
         mbar_expect_tx(bar, (uint32_t)(NEB * C::NQD * Q3) * 8u);
#pragma unroll
         for (int e = 0; e < NEB; e++)
         {
            bulk_g2s(qd + e * C::DS, qslab + (int64_t)e * (C::NQD * Q3), (uint32_t)(C::NQD * Q3) * 8u, bar);
         }
      }
      else
      {
         mbar_expect_tx(bar, (uint32_t)(nev * C::NQD * Q3) * 8u);
         for (int e = 0; e < nev; e++)
         {
            bulk_g2s(qd + e * C::DS, qslab + (int64_t)e * (C::NQD * Q3), (uint32_t)(C::NQD * Q3) * 8u, bar);
         }
      }
   }
   if (pf > 0 && t == ISSUER + 1)
   {
      const int64_t ep = e0 + (int64_t)pf * NEB;
      if (ep + NEB <= ne)
      {
         const uintptr_t q = reinterpret_cast<uintptr_t>(qdata + ep * (int64_t)(C::NQD * Q3)) & ~(uintptr_t)15;
         l2_prefetch_bulk(reinterpret_cast<const void *>(q), ((uint32_t)(NEB * C::NQD * Q3) * 8u) & ~15u);
         const uintptr_t a = reinterpret_cast<uintptr_t>(idx + ep * D3) & ~(uintptr_t)15;
         l2_prefetch_bulk(reinterpret_cast<const void *>(a), ((uint32_t)(NEB * D3) * 4u) & ~15u);
      }
   }
   if (has1)
   {
      double u[D];
#pragma unroll
      for (int dy = 0; dy < D; dy++) { u[dy] = gi[dy] >= 0 ? __ldg(x + gi[dy]) : 0.0; }
      S::stage1(bs, u, r1 + (lc & 0xffffu)); // absent elements of a ragged tile: zeros
   }
   __syncthreads();
   if (t < C::T2) { S::stage2(bs, r1, r2, t); }
   __syncthreads();
   if (t < C::T3) { S::stage3(bs, r2, qd, t, bar, 0); }
   __syncthreads();
   if (t < C::T2) { S::stage4(bs, r2, r1, t); }
   __syncthreads();
   if (act1)
   {
      double v[D];
      S::stage5(bs, r1 + (lc & 0xffffu), v);
#pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         if (gi[dy] >= 0) { atomicAdd(y + gi[dy], v[dy]); }
      }
   }
}

template <int KIND, int D, int Q, int NEB, int MINB>
int launch_3d_eo_lean(b2f_pa *pa, const double *x, double *y, const int *idx)
{
   using C = PipeCfg<KIND, D, Q, NEB>;
   constexpr int NTAB = ((C::T1 + 3) / 4) * 4;
   static_assert(D * C::SA * C::NA < 65536 && C::D3 <= 1024 && NEB <= 32, "lane word packing");
   b2f_ctx *ctx = pa->ctx;
   if (!idx || pa->vd > 1 || pa->collocated)
   {
      set_error("pa_apply_3d_eo_lean: fused scalar, non-collocated form only");
      return B2F_ERR_UNSUPPORTED;
   }
   if ((int)C::SMEM > ctx->max_smem_optin)
   {
      set_error("pa_apply_3d_eo_lean<D=%d,Q=%d,NEB=%d> needs %zu B of shared memory", D, Q, NEB, C::SMEM);
      return B2F_ERR_UNSUPPORTED;
   }
   EOBasis<D, Q> bs;
   if (!fill_eo_basis(pa, bs))
   {
      set_error("pa_apply_3d_eo_lean: the basis tables are not symmetric");
      return B2F_ERR_UNSUPPORTED;
   }
   LaneTab<NTAB> lt;
   for (int t = 0; t < NTAB; t++)
   {
      // eo_line_coords on the host (same lane orders)
      int e1, dx1, dz1;
      if (C::L.order == 0) { e1 = t % NEB; dx1 = (t / NEB) % D; dz1 = t / (NEB * D); }
      else if (C::L.order == 1) { dx1 = t % D; dz1 = (t / D) % D; e1 = t / (D * D); }
      else if (C::L.order == 2) { e1 = t % NEB; dz1 = (t / NEB) % D; dx1 = t / (NEB * D); }
      else { dz1 = t % D; e1 = (t / D) % NEB; dx1 = t / (NEB * D); }
      if (t >= C::T1) { lt.v[t] = 0; continue; }
      const unsigned a1 = (unsigned)(Q * e1 + C::YQ * dx1 + C::SA * dz1);
      const unsigned loc1 = (unsigned)(dx1 + C::D2 * dz1);
      lt.v[t] = a1 | (loc1 << 16) | ((unsigned)e1 << 26);
   }
   const int64_t e_first = pa->rb, e_last = pa->re < 0 ? pa->ne : pa->re;
   if (e_last <= e_first) { return B2F_OK; }
   const int64_t grid = cdiv(e_last - e_first, NEB);
   auto kern = pa_apply_3d_eo_lean_kernel<KIND, D, Q, NEB, MINB>;
   static int resident = 0;
   if (resident == 0)
   {
      B2F_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
      int occ = 0;
      B2F_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, C::SMEM));
      resident = (occ < 1 ? 1 : occ) * ctx->sm_count;
   }
   const char *pfs = getenv("B2F_PF");
   const int pf = pfs ? atoi(pfs) : resident;
   timing_mark(ctx);
   kern<<<(unsigned)grid, C::NT, C::SMEM, ctx->stream>>>(bs, lt, pa->data, x, y, idx, e_last, pf, e_first);
   timing_mark(ctx);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

} // namespace b2f
