// mfem_b200/csrc/pa_apply_pipe.cuh -- the 3D hot loop: persistent, software-pipelined,
// sum-factorised PA operator kernel (diffusion and mass, fused L-vector or E-vector form).
//
// Replaces MFEM's SmemPADiffusionApply3D / PADiffusionApply3D
// (fem/integ/bilininteg_diffusion_kernels.hpp:796-1215), SmemPAMassApply3D / PAMassApply3D
// (fem/integ/bilininteg_mass_kernels.hpp:664-1165) and -- FUSED -- ElementRestriction::Mult /
// MultTranspose (fem/restriction.cpp:116,159) plus the E-vector zero fill
// (fem/bilinearform_ext.cpp:543-557).
//
// Why it looks the way it does (measured on B200, profiles/r01_*):
//  * The kernel is bound by HBM (11.9 KB of quadrature data per p=4 element, used once), but
//    a CTA that loads its slab, waits, and then computes keeps too few bytes in flight
//    (Little's law: ~45 KB per SM must be in flight all the time).  So CTAs are PERSISTENT
//    and every stream is prefetched one tile ahead while the current tile computes:
//      - quadrature data: one TMA bulk copy per element (cp.async.bulk + mbarrier) into a
//        padded shared-memory slab, issued as soon as the D-apply stage of the previous
//        tile has released the slab;
//      - gather indices: cp.async (LDGSTS) into shared memory one tile ahead;
//      - the L-vector values x[idx]: plain loads into registers, issued half a tile ahead.
//  * "Pencil" decomposition: in every contraction stage one thread owns one 1D line, loads
//    it into registers once and produces every output of every field it feeds, so each
//    shared-memory word is read once per stage and B, G enter the DFMAs as constant-bank
//    operands (they are kernel parameters).
//  * Stage order y -> z -> x with yq = qy + Q*e as the FASTEST thread and address index of
//    stages 2-4: all inter-stage buffers are read and written as consecutive addresses by
//    consecutive lanes (no bank conflicts, no transposition), the x-stage can overwrite its
//    input in place (each thread owns its column: no barrier, no staging registers), and a
//    lane's quadrature-data rows are Q contiguous doubles read with 128-bit LDS.
//  * FUSED: y is scattered with FP64 reductions in L2 (RED.E.ADD.F64); index -1 marks an
//    essential dof: it reads 0 and is not scattered (ConstrainedOperator, operator.cpp:601).
#pragma once
#include "pa_apply.cuh"
#include "pa_layout_table.h"

namespace b2f
{

// stage-1/5 lane order, slab-stride residue and quadrature-slab padding: generated offline by
// tools/gen_layout_table.py (simulation of every shared-memory access pattern)
struct PipeLayout
{
   int order; // 0: (e, dx, dz) e fastest   1: (dx, dz, e)   2: (e, dz, dx)   3: (dz, e, dx)
   int ra;    // SA % 16
   int dpad;  // doubles appended to each element's quadrature-data slab
};

__host__ __device__ constexpr PipeLayout pipe_layout(int kind, int D, int Q, int NEB)
{
   if (D < 2 || D > 9 || Q < D || Q > D + 2 || NEB < 1 || NEB > 32) { return {0, 0, 0}; }
   const LayoutEntry e = kLayoutTable[kind == B2F_KIND_DIFFUSION ? 0 : 1][D - 2][Q - D][NEB - 1];
   return {e.order, e.ra, e.dpad};
}

template <int KIND, int D, int Q, int NEB>
struct PipeCfg
{
   static constexpr bool DIFF = (KIND == B2F_KIND_DIFFUSION);
   static constexpr int NA = DIFF ? 2 : 1;  // fields between the y and z stages
   static constexpr int NC = DIFF ? 3 : 1;  // fields between the z and x stages
   static constexpr int NQD = DIFF ? 6 : 1; // quadrature-data components (symmetric)
   static constexpr int D2 = D * D, D3 = D * D * D, Q2 = Q * Q, Q3 = Q * Q * Q;
   static constexpr int YQ = Q * NEB;   // yq = qy + Q*e
   static constexpr int T1 = NEB * D2;  // y, y^T tasks (dx, dz, e)
   static constexpr int T2 = D * YQ;    // z, z^T tasks (yq, dx)
   static constexpr int T3 = Q * YQ;    // x / D / x^T tasks (yq, qz)
   static constexpr int NT = ((T3 + 31) / 32) * 32;
   static constexpr PipeLayout L = pipe_layout(KIND, D, Q, NEB);
   static constexpr int SA = pad16(D * YQ, L.ra); // A,F[f][dz][dx][yq]
   static constexpr int SC = pad16(Q * YQ, YQ);   // C,E[f][dx][qz][yq]
   // stage 1 / 5: a line's Q values as Q/2 128-bit words (line starts are even: Q, YQ, SA even)
   static constexpr bool VEC1 = (Q % 2 == 0) && (SA % 2 == 0);
   static constexpr int R1 = NA * D * SA;
   static constexpr int R2 = NC * D * SC;
   // quadrature data: per-element slabs, padded so that 128-bit row reads by consecutive
   // lanes (qy fastest, then e) hit distinct bank groups across element boundaries
   static constexpr bool TMA_OK = ((NQD * Q3) % 2) == 0; // slab bytes multiple of 16
   static constexpr bool VEC = TMA_OK && (Q % 2 == 0);   // rows 16-byte aligned
   static constexpr int DS = NQD * Q3 + (TMA_OK ? L.dpad : 0);
   static constexpr int QD = NEB * DS;
   static constexpr int NIDX = ((NEB * D3 + 3) / 4) * 4; // ints
   static constexpr size_t SMEM = 16 + sizeof(double) * (size_t)(QD + R1 + R2) + sizeof(int) * NIDX;
   static_assert(T1 <= NT && T2 <= NT, "one task per thread per stage");
};

__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
   asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
   asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int KIND, int D, int Q, int NEB, bool FUSED, bool COLLOC, int MINB>
__global__ void __launch_bounds__(PipeCfg<KIND, D, Q, NEB>::NT, MINB)
pa_apply_3d_pipe_kernel(const __grid_constant__ Basis<D, Q> bs, const double *__restrict__ qdata,
                        const double *__restrict__ x, double *__restrict__ y,
                        const int *__restrict__ idx, const int64_t ne)
{
   using C = PipeCfg<KIND, D, Q, NEB>;
   constexpr bool DIFF = C::DIFF;
   constexpr int D2 = C::D2, D3 = C::D3, Q2 = C::Q2, Q3 = C::Q3, YQ = C::YQ;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
   double *qd = reinterpret_cast<double *>(smem_raw + 16);
   double *r1 = qd + C::QD;
   double *r2 = r1 + C::R1;
   int *idx_s = reinterpret_cast<int *>(r2 + C::R2);

   const int t = threadIdx.x;
   const int64_t ntiles = (ne + NEB - 1) / NEB;

   // ---- per-thread task coordinates ------------------------------------------------------
   // stages 1 and 5: (dx, dz, e)
   int e1, dx1, dz1;
   if (C::L.order == 0) { e1 = t % NEB; dx1 = (t / NEB) % D; dz1 = t / (NEB * D); }
   else if (C::L.order == 1) { dx1 = t % D; dz1 = (t / D) % D; e1 = t / D2; }
   else if (C::L.order == 2) { e1 = t % NEB; dz1 = (t / NEB) % D; dx1 = t / (NEB * D); }
   else { dz1 = t % D; e1 = (t / D) % NEB; dx1 = t / (NEB * D); }
   const bool has1 = t < C::T1;
   const int loc1 = dx1 + D2 * dz1;                      // + D*dy : position inside the element
   const int a1 = Q * e1 + YQ * dx1 + C::SA * dz1;       // + qy   : A/F address of this line
   // stage 3: (yq, qz)
   const int yq3 = t % YQ, qz3 = t / YQ;
   const int e3 = yq3 / Q, qy3 = yq3 - e3 * Q;

   auto issue_qdata = [&](int64_t tile)
   {
      const int64_t e0 = tile * NEB;
      const int nev = (int)((ne - e0) < NEB ? (ne - e0) : NEB);
      const double *src = qdata + e0 * (int64_t)(C::NQD * Q3);
      if constexpr (C::TMA_OK)
      {
         if (t == 0)
         {
            fence_proxy_async(); // the slab was just read through the generic proxy
            mbar_expect_tx(bar, (uint32_t)(nev * C::NQD * Q3) * 8u);
            for (int e = 0; e < nev; e++)
            {
               bulk_g2s(qd + e * C::DS, src + (int64_t)e * (C::NQD * Q3), (uint32_t)(C::NQD * Q3) * 8u, bar);
            }
         }
      }
      else
      {
         for (int i = t; i < nev * C::NQD * Q3; i += C::NT) { cp_async8(qd + i, src + i); }
      }
   };
   auto issue_idx = [&](int64_t tile)
   {
      if constexpr (FUSED)
      {
         const int64_t e0 = tile * NEB;
         const int nev = (int)((ne - e0) < NEB ? (ne - e0) : NEB);
         const int *src = idx + e0 * D3;
         for (int i = t; i < nev * D3; i += C::NT) { cp_async4(idx_s + i, src + i); }
      }
   };
   // gather of one tile's inputs into registers (indices come from idx_s)
   auto load_inputs = [&](int64_t tile, int (&gi)[D], double (&u)[D])
   {
      const int64_t e0 = tile * NEB;
      const int nev = (int)((ne - e0) < NEB ? (ne - e0) : NEB);
#pragma unroll
      for (int dy = 0; dy < D; dy++) { gi[dy] = -1; u[dy] = 0.0; }
      if (has1 && e1 < nev)
      {
         if constexpr (FUSED)
         {
#pragma unroll
            for (int dy = 0; dy < D; dy++) { gi[dy] = idx_s[e1 * D3 + loc1 + D * dy]; }
#pragma unroll
            for (int dy = 0; dy < D; dy++) { u[dy] = gi[dy] >= 0 ? __ldg(x + gi[dy]) : 0.0; }
         }
         else
         {
            const double *xe = x + (e0 + e1) * D3 + loc1;
#pragma unroll
            for (int dy = 0; dy < D; dy++) { gi[dy] = 0; u[dy] = __ldg(xe + D * dy); }
         }
      }
   };

   // ---- prologue: first tile's streams -------------------------------------------------
   int64_t tile = blockIdx.x;
   if (tile >= ntiles) { return; }
   if (t == 0) { mbar_init(bar, 1); }
   __syncthreads();
   issue_qdata(tile);
   issue_idx(tile);
   cp_async_wait_all();
   __syncthreads();
   int gi[D], gi_n[D];
   double u[D], u_n[D];
   load_inputs(tile, gi, u);
   uint32_t parity = 0;

   for (; tile < ntiles; tile += gridDim.x)
   {
      const int64_t e0 = tile * NEB;
      const int nev = (int)((ne - e0) < NEB ? (ne - e0) : NEB);
      const int64_t next = tile + gridDim.x;
      const bool more = next < ntiles;

      // ---- stage 1 (y): line (dx, dz, e): u[dy] -> B u, G u at qy ----------------------
      if (has1)
      {
         double *A0 = r1 + a1;
#pragma unroll
         for (int qy = 0; qy < Q; qy++)
         {
            double vb = 0.0, vg = 0.0;
            if (COLLOC) { vb = u[qy < D ? qy : 0]; }
#pragma unroll
            for (int dy = 0; dy < D; dy++)
            {
               if (!COLLOC) { vb = fma(bs.B[qy][dy], u[dy], vb); }
               if (DIFF) { vg = fma(bs.G[qy][dy], u[dy], vg); }
            }
            A0[qy] = vb;
            if (DIFF) { A0[D * C::SA + qy] = vg; }
         }
      }
      if (more) { issue_idx(next); } // idx_s was consumed when this tile's inputs were gathered
      __syncthreads();

      // ---- stage 2 (z): line t2 = (yq, dx): contraction over dz -------------------------
      if (t < C::T2)
      {
         double a0[D], a1v[D];
#pragma unroll
         for (int dz = 0; dz < D; dz++)
         {
            a0[dz] = r1[t + C::SA * dz];
            if (DIFF) { a1v[dz] = r1[t + C::SA * (D + dz)]; }
         }
         const int yq = t % YQ, dx = t / YQ;
         double *C0 = r2 + yq + C::SC * dx;
#pragma unroll
         for (int qz = 0; qz < Q; qz++)
         {
            double bb = 0.0, gyv = 0.0, gzv = 0.0;
            if (COLLOC) { bb = a0[qz < D ? qz : 0]; if (DIFF) { gyv = a1v[qz < D ? qz : 0]; } }
#pragma unroll
            for (int dz = 0; dz < D; dz++)
            {
               if (!COLLOC)
               {
                  bb = fma(bs.B[qz][dz], a0[dz], bb);
                  if (DIFF) { gyv = fma(bs.B[qz][dz], a1v[dz], gyv); }
               }
               if (DIFF) { gzv = fma(bs.G[qz][dz], a0[dz], gzv); }
            }
            C0[YQ * qz] = bb;
            if (DIFF)
            {
               C0[D * C::SC + YQ * qz] = gyv;     // d/dy
               C0[2 * D * C::SC + YQ * qz] = gzv; // d/dz
            }
         }
      }
      if constexpr (!C::TMA_OK) { cp_async_wait_all(); }
      __syncthreads();

      // ---- stage 3 (x, D, x^T) in place: line t3 = (yq, qz) -----------------------------
      if (t < C::T3)
      {
         double *C0 = r2 + yq3 + YQ * qz3;
         double c0[D], c1[D], c2[D];
#pragma unroll
         for (int dx = 0; dx < D; dx++)
         {
            c0[dx] = C0[C::SC * dx];
            if (DIFF)
            {
               c1[dx] = C0[C::SC * (D + dx)];
               c2[dx] = C0[C::SC * (2 * D + dx)];
            }
         }
         double gx[Q], gy[Q], gz[Q];
#pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            double vx = 0.0, vy = 0.0, vz = 0.0;
            if (COLLOC)
            {
               if (DIFF) { vy = c1[qx < D ? qx : 0]; vz = c2[qx < D ? qx : 0]; }
               else { vx = c0[qx < D ? qx : 0]; }
            }
#pragma unroll
            for (int dx = 0; dx < D; dx++)
            {
               if (DIFF)
               {
                  vx = fma(bs.G[qx][dx], c0[dx], vx);
                  if (!COLLOC)
                  {
                     vy = fma(bs.B[qx][dx], c1[dx], vy);
                     vz = fma(bs.B[qx][dx], c2[dx], vz);
                  }
               }
               else if (!COLLOC) { vx = fma(bs.B[qx][dx], c0[dx], vx); }
            }
            gx[qx] = vx; gy[qx] = vy; gz[qx] = vz;
         }
         // quadrature data of this line: NQD rows of Q doubles from the (TMA-filled) slab
         if constexpr (C::TMA_OK) { mbar_wait(bar, parity); }
         const double *dq = qd + e3 * C::DS + Q * (qy3 + Q * qz3);
         double dv[C::NQD][Q];
#pragma unroll
         for (int k = 0; k < C::NQD; k++)
         {
            if constexpr (C::VEC) { load_row<Q, false>(dq + k * Q3, dv[k]); }
            else
            {
#pragma unroll
               for (int qx = 0; qx < Q; qx++) { dv[k][qx] = dq[k * Q3 + qx]; }
            }
         }
         if (DIFF)
         {
#pragma unroll
            for (int qx = 0; qx < Q; qx++)
            {
               const double O11 = dv[0][qx], O12 = dv[1 % C::NQD][qx], O13 = dv[2 % C::NQD][qx];
               const double O22 = dv[3 % C::NQD][qx], O23 = dv[4 % C::NQD][qx], O33 = dv[5 % C::NQD][qx];
               const double ux = gx[qx], uy = gy[qx], uz = gz[qx];
               gx[qx] = (O11 * ux) + (O12 * uy) + (O13 * uz);
               gy[qx] = (O12 * ux) + (O22 * uy) + (O23 * uz);
               gz[qx] = (O13 * ux) + (O23 * uy) + (O33 * uz);
            }
         }
         else
         {
#pragma unroll
            for (int qx = 0; qx < Q; qx++) { gx[qx] *= dv[0][qx]; }
         }
         // x^T, written over this thread's own input column
#pragma unroll
         for (int dx = 0; dx < D; dx++)
         {
            double va = 0.0, vb = 0.0, vc = 0.0;
            if (COLLOC)
            {
               if (DIFF) { vb = gy[dx]; vc = gz[dx]; }
               else { va = gx[dx]; }
            }
#pragma unroll
            for (int qx = 0; qx < Q; qx++)
            {
               if (DIFF)
               {
                  va = fma(bs.G[qx][dx], gx[qx], va);
                  if (!COLLOC)
                  {
                     vb = fma(bs.B[qx][dx], gy[qx], vb);
                     vc = fma(bs.B[qx][dx], gz[qx], vc);
                  }
               }
               else if (!COLLOC) { va = fma(bs.B[qx][dx], gx[qx], va); }
            }
            C0[C::SC * dx] = va;
            if (DIFF)
            {
               C0[C::SC * (D + dx)] = vb;
               C0[C::SC * (2 * D + dx)] = vc;
            }
         }
      }
      parity ^= 1u;
      if (FUSED && more) { cp_async_wait_all(); } // next tile's indices have landed (this thread's)
      __syncthreads();
      // the slab is free: stream the next tile's quadrature data while stages 4, 5 and the
      // next tile's stages 1, 2 run; gather the next tile's inputs into registers
      if (more)
      {
         issue_qdata(next);
         load_inputs(next, gi_n, u_n);
      }

      // ---- stage 4 (z^T): line t2 = (yq, dx) --------------------------------------------
      if (t < C::T2)
      {
         const int yq = t % YQ, dx = t / YQ;
         const double *E0 = r2 + yq + C::SC * dx;
         double ea[Q], eb[Q], ec[Q];
#pragma unroll
         for (int qz = 0; qz < Q; qz++)
         {
            ea[qz] = E0[YQ * qz];
            if (DIFF)
            {
               eb[qz] = E0[D * C::SC + YQ * qz];
               ec[qz] = E0[2 * D * C::SC + YQ * qz];
            }
         }
#pragma unroll
         for (int dz = 0; dz < D; dz++)
         {
            double p0 = 0.0, p1 = 0.0;
            if (COLLOC) { p0 = ea[dz]; if (DIFF) { p1 = eb[dz]; } }
#pragma unroll
            for (int qz = 0; qz < Q; qz++)
            {
               if (!COLLOC)
               {
                  p0 = fma(bs.B[qz][dz], ea[qz], p0);
                  if (DIFF) { p1 = fma(bs.B[qz][dz], eb[qz], p1); }
               }
               if (DIFF) { p0 = fma(bs.G[qz][dz], ec[qz], p0); }
            }
            r1[t + C::SA * dz] = p0;
            if (DIFF) { r1[t + C::SA * (D + dz)] = p1; }
         }
      }
      __syncthreads();

      // ---- stage 5 (y^T) + scatter: line (dx, dz, e) --------------------------------------
      if (has1 && e1 < nev)
      {
         const double *F0 = r1 + a1;
         double f0[Q], f1[Q];
#pragma unroll
         for (int qy = 0; qy < Q; qy++)
         {
            f0[qy] = F0[qy];
            if (DIFF) { f1[qy] = F0[D * C::SA + qy]; }
         }
         double *ye = y + (e0 + e1) * D3 + loc1; // E-vector form
#pragma unroll
         for (int dy = 0; dy < D; dy++)
         {
            double v = 0.0;
            if (COLLOC) { v = f0[dy]; }
#pragma unroll
            for (int qy = 0; qy < Q; qy++)
            {
               if (!COLLOC) { v = fma(bs.B[qy][dy], f0[qy], v); }
               if (DIFF) { v = fma(bs.G[qy][dy], f1[qy], v); }
            }
            if (FUSED)
            {
               if (gi[dy] >= 0) { atomicAdd(y + gi[dy], v); }
            }
            else { ye[D * dy] += v; }
         }
      }
      if (more)
      {
#pragma unroll
         for (int dy = 0; dy < D; dy++) { gi[dy] = gi_n[dy]; u[dy] = u_n[dy]; }
      }
      __syncthreads(); // F consumed before the next tile's stage 1 overwrites the region
   }
}

template <int KIND, int D, int Q, int NEB, int MINB = 1>
int launch_3d_pipe(b2f_pa *pa, const double *x, double *y, const int *idx)
{
   using C = PipeCfg<KIND, D, Q, NEB>;
   b2f_ctx *ctx = pa->ctx;
   if ((int)C::SMEM > ctx->max_smem_optin)
   {
      set_error("pa_apply_3d_pipe<D=%d,Q=%d,NEB=%d> needs %zu B of shared memory", D, Q, NEB, C::SMEM);
      return B2F_ERR_UNSUPPORTED;
   }
   Basis<D, Q> bs;
   fill_basis(pa, bs);
   const int64_t ntiles = cdiv(pa->ne, NEB);
   auto go = [&](auto kern) -> int
   {
      B2F_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
      int occ = 0;
      B2F_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, C::SMEM));
      if (occ < 1) { occ = 1; }
      int64_t grid = (int64_t)occ * ctx->sm_count;
      if (grid > ntiles) { grid = ntiles; }
      timing_mark(ctx);
      kern<<<(unsigned)grid, C::NT, C::SMEM, ctx->stream>>>(bs, pa->data, x, y, idx, pa->ne);
      timing_mark(ctx);
      B2F_LAUNCH_CHECK(ctx);
      return B2F_OK;
   };
   const bool fused = idx != nullptr;
   if (pa->collocated && D == Q)
   {
      constexpr bool CL = (D == Q);
      return fused ? go(pa_apply_3d_pipe_kernel<KIND, D, Q, NEB, true, CL, MINB>)
             : go(pa_apply_3d_pipe_kernel<KIND, D, Q, NEB, false, CL, MINB>);
   }
   return fused ? go(pa_apply_3d_pipe_kernel<KIND, D, Q, NEB, true, false, MINB>)
          : go(pa_apply_3d_pipe_kernel<KIND, D, Q, NEB, false, false, MINB>);
}

} // namespace b2f
