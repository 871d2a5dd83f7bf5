// mfem_b200/csrc/pa_apply.cuh -- the hot loop: sum-factorised PA operator kernels.
//
// Replaces MFEM's (Smem)PADiffusionApply2D/3D (fem/integ/bilininteg_diffusion_kernels.hpp:
// 514-1215) and (Smem)PAMassApply2D/3D (fem/integ/bilininteg_mass_kernels.hpp:440-1165),
// and -- in FUSED mode -- also ElementRestriction::Mult / MultTranspose
// (fem/restriction.cpp:116,159) and the zero-fill of the E-vector
// (fem/bilinearform_ext.cpp:543-557), so the E-vectors never exist in HBM.
//
// Design (B200 / sm_100a; FP64 FMA pipes, not tensor cores -- the contractions are
// (p+1)x(p+2) matrices):
//  * One CTA owns NEB consecutive elements.  The quadrature data D of those elements is
//    one contiguous slab in HBM ((Q^3, ncomp, NE) layout); a single elected thread pulls it
//    into shared memory with ONE bulk-async copy (TMA, cp.async.bulk + mbarrier
//    complete_tx) at kernel entry, so the dominant HBM stream is in flight while the CTA
//    does the gather and the first two contraction stages, and it costs no registers.
//  * "Pencil" decomposition: in every contraction stage one thread owns one 1D line of
//    the tensor, loads its D1D (or Q1D) inputs into registers once and produces all Q1D
//    (or D1D) outputs for every field it feeds (value + derivative), so each shared-memory
//    word is read exactly once per stage: 2*(2D^2Q + 3DQ^2) doubles of smem traffic per
//    element instead of one load per FMA.  The basis matrices B, G are kernel parameters
//    (constant bank) and enter the DFMAs as immediate constant operands.
//  * Stage order z -> y -> x so that (a) the L-vector gather / scatter indices are read
//    coalesced (E-index is x-fastest), and (b) the D-apply stage owns x-lines, i.e. reads
//    D rows of Q1D contiguous doubles.
//  * Inter-stage buffers are laid out [contraction index][consumer task id], which makes
//    every read conflict-free; slab strides are padded (mod 16 doubles) for the writes.
//  * FUSED: x is gathered straight into registers (index -1 = essential dof: reads 0 and
//    is not scattered), y is scattered with FP64 atomics (RED.ADD.F64 in L2).
#pragma once
#include <map>
#include <cstdlib>

#include "common.cuh"

namespace b2f
{

template <int D, int Q>
struct Basis
{
   double B[Q][D];
   double G[Q][D];
};

__host__ __device__ constexpr int pad16(int n, int residue)
{
   // smallest m >= n with m % 16 == residue % 16
   int r = ((residue % 16) + 16) % 16;
   int m = n - (n % 16) + r;
   return m >= n ? m : m + 16;
}

// ---- bulk-async (TMA) helpers -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
   return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                "r"(bytes)
                : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                         uint64_t *bar)
{
   asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
      "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
   uint32_t ok = 0;
   while (!ok)
   {
      asm volatile(
         "{\n\t.reg .pred p;\n\t"
         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
         "selp.u32 %0, 1, 0, p;\n\t}"
         : "=r"(ok)
         : "r"(smem_u32(bar)), "r"(parity)
         : "memory");
   }
}

// Pull `n` doubles of quadrature data into shared memory.  One TMA bulk copy when the
// slab is 16-byte aligned and sized (always true except for odd tails of odd-Q mass
// data), otherwise a cooperative copy.  Returns true when the mbarrier must be waited on.
template <int NT>
__device__ __forceinline__ bool load_qdata(double *dst, const double *src, int n, uint64_t *bar)
{
   const bool tma_ok = ((n & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
   if (tma_ok)
   {
      if (threadIdx.x == 0)
      {
         mbar_init(bar, 1);
         mbar_expect_tx(bar, (uint32_t)n * 8u);
         bulk_g2s(dst, src, (uint32_t)n * 8u, bar);
      }
   }
   else
   {
      for (int i = threadIdx.x; i < n; i += NT) { dst[i] = __ldg(src + i); }
   }
   return tma_ok;
}

// =====================================================================================
// 3D kernel
// =====================================================================================
// DMODE: how the quadrature data reaches the D-apply stage
//   0  TMA bulk copy of the CTA's slab into shared memory, scalar LDS reads
//   1  same, rows read back with 128-bit LDS (conflict-free for even Q)
//   2  no staging: L2 bulk prefetch of the slab at kernel entry, rows loaded straight into
//      registers with 128-bit LDG right before use (each D value is used exactly once)
//   3  as 2, loads issued before the y-stage so they overlap it (costs registers)
//   4  as 2 without the L2 prefetch
template <int KIND, int D, int Q, int NEB, int DMODE = 0>
struct Cfg3D
{
   static constexpr bool DIFF = (KIND == B2F_KIND_DIFFUSION);
   static constexpr int NA = DIFF ? 2 : 1;  // fields after the z / before the z^T stage
   static constexpr int NC = DIFF ? 3 : 1;  // fields after the y / before the y^T stage
   static constexpr int NQD = DIFF ? 6 : 1; // quadrature-data components (symmetric)
   static constexpr int D2 = D * D, D3 = D * D * D, Q2 = Q * Q, Q3 = Q * Q * Q;
   static constexpr int T1 = NEB * D2;    // z and z^T stage tasks   (dx,dy,e)
   static constexpr int T2 = NEB * Q * D; // y and y^T stage tasks   (dx,qz,e)
   static constexpr int T3 = NEB * Q2;    // x / D / x^T stage tasks (qy,qz,e)
   static constexpr int NT = ((T3 + 31) / 32) * 32;
   // slab strides (doubles), padded for conflict-poor writes
   static constexpr int SA = pad16(T2, D); // A[field][dy][t2]
   static constexpr int SF = pad16(T1, D); // F[field][qz][t1]
   static constexpr int SC = pad16(T3, 1); // C[field][dx][t3]
   static constexpr int SE = pad16(T2, 1); // E[field][qy][t2]
   static constexpr int R1 = (NA * D * SA > NA * Q * SF) ? NA * D * SA : NA * Q * SF;
   static constexpr int R2 = (NC * D * SC > NC * Q * SE) ? NC * D * SC : NC * Q * SE;
   static constexpr bool STAGED = DMODE <= 1;                // D goes through shared memory
   static constexpr int QD = STAGED ? NEB * NQD * Q3 : 0;    // quadrature data tile
   static constexpr size_t SMEM = 16 + sizeof(double) * (size_t)(QD + R1 + R2);
   static_assert(T1 <= NT && T2 <= NT, "one task per thread per stage");
};

// one row of Q consecutive quadrature-data values -> registers (128-bit accesses when the
// row is 16-byte aligned, i.e. Q even; works for shared and global pointers alike)
template <int Q, bool GLOBAL>
__device__ __forceinline__ void load_row(const double *__restrict__ p, double (&v)[Q])
{
   if constexpr (Q % 2 == 0)
   {
      const double2 *p2 = reinterpret_cast<const double2 *>(p);
#pragma unroll
      for (int i = 0; i < Q / 2; i++)
      {
         const double2 w = GLOBAL ? __ldg(p2 + i) : p2[i];
         v[2 * i] = w.x; v[2 * i + 1] = w.y;
      }
   }
   else
   {
#pragma unroll
      for (int i = 0; i < Q; i++) { v[i] = GLOBAL ? __ldg(p + i) : p[i]; }
   }
}

__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes)
{
   asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int KIND, int D, int Q, int NEB, bool FUSED, bool COLLOC, int DMODE, int MINB>
__global__ void __launch_bounds__(Cfg3D<KIND, D, Q, NEB, DMODE>::NT, MINB)
pa_apply_3d_kernel(const __grid_constant__ Basis<D, Q> bs, const double *__restrict__ qdata,
                   const double *__restrict__ x, double *__restrict__ y,
                   const int *__restrict__ idx, const int64_t ne, const int pf)
{
   using C = Cfg3D<KIND, D, Q, NEB, DMODE>;
   constexpr bool DIFF = C::DIFF;
   constexpr int D2 = C::D2, D3 = C::D3, Q2 = C::Q2, Q3 = C::Q3;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
   double *qd = reinterpret_cast<double *>(smem_raw + 16);
   double *r1 = qd + C::QD;
   double *r2 = r1 + C::R1;

   const int t = threadIdx.x;
   const int64_t e0 = (int64_t)blockIdx.x * NEB;
   const int nev = (int)((ne - e0) < NEB ? (ne - e0) : NEB); // valid elements in this tile

   // ---- kick off the quadrature-data stream (TMA into smem, or an L2 prefetch) -------
   bool use_bar = false;
   const double *qslab = qdata + e0 * (int64_t)(C::NQD * Q3);
   if constexpr (C::STAGED)
   {
      use_bar = load_qdata<C::NT>(qd, qslab, nev * C::NQD * Q3, bar);
      // keep HBM busy while this CTA computes: pull the slab (and gather indices) of the CTA
      // that will run `pf` blocks later into L2, so its TMA copy is an L2 hit
      if (pf > 0 && t == 32)
      {
         const int64_t ep = e0 + (int64_t)pf * NEB;
         if (ep + NEB <= ne)
         {
            const uintptr_t q = reinterpret_cast<uintptr_t>(qdata + ep * (int64_t)(C::NQD * Q3)) & ~(uintptr_t)15;
            l2_prefetch_bulk(reinterpret_cast<const void *>(q), ((uint32_t)(NEB * C::NQD * Q3) * 8u) & ~15u);
            if (FUSED)
            {
               const uintptr_t a = reinterpret_cast<uintptr_t>(idx + ep * D3) & ~(uintptr_t)15;
               l2_prefetch_bulk(reinterpret_cast<const void *>(a), ((uint32_t)(NEB * D3) * 4u) & ~15u);
            }
         }
      }
   }
   else if constexpr (DMODE != 4)
   {
      if (t == 0)
      {
         const uintptr_t q = reinterpret_cast<uintptr_t>(qslab) & ~(uintptr_t)15;
         l2_prefetch_bulk(reinterpret_cast<const void *>(q), ((uint32_t)(nev * C::NQD * Q3) * 8u) & ~15u);
      }
   }

   // ---- stage 1 (z): gather + contraction over dz ; task (dx,dy,e) ------------------
   int gi[D]; // gather indices, reused by the scatter
   {
      const int e = t / D2, r = t - e * D2;
      const bool active = (t < C::T1) && (e < nev);
      double u[D];
#pragma unroll
      for (int dz = 0; dz < D; dz++) { gi[dz] = -1; u[dz] = 0.0; }
      if (active)
      {
         const int64_t base = (e0 + e) * D3 + r;
         if (FUSED)
         {
#pragma unroll
            for (int dz = 0; dz < D; dz++) { gi[dz] = __ldg(idx + base + dz * D2); }
#pragma unroll
            for (int dz = 0; dz < D; dz++) { u[dz] = gi[dz] >= 0 ? __ldg(x + gi[dz]) : 0.0; }
         }
         else
         {
#pragma unroll
            for (int dz = 0; dz < D; dz++) { u[dz] = __ldg(x + base + dz * D2); }
         }
      }
      if (t < C::T1)
      {
         const int dy = r / D, dx = r - dy * D;
         double *A0 = r1 + dy * C::SA + dx + D * (Q * e);
#pragma unroll
         for (int qz = 0; qz < Q; qz++)
         {
            double vb = 0.0, vg = 0.0;
            if (COLLOC) { vb = u[qz < D ? qz : 0]; }
#pragma unroll
            for (int dz = 0; dz < D; dz++)
            {
               if (!COLLOC) { vb = fma(bs.B[qz][dz], u[dz], vb); }
               if (DIFF) { vg = fma(bs.G[qz][dz], u[dz], vg); }
            }
            A0[D * qz] = vb;
            if (DIFF) { A0[D * C::SA + D * qz] = vg; }
         }
      }
   }
   __syncthreads();

   // DMODE 3: issue this thread's quadrature-data loads now; they land during the y-stage
   double dearly[DMODE == 3 ? C::NQD : 1][DMODE == 3 ? Q : 1];
   if constexpr (DMODE == 3)
   {
      const int e = t / Q2, row = t - e * Q2;
#pragma unroll
      for (int k = 0; k < C::NQD; k++)
#pragma unroll
         for (int qx = 0; qx < Q; qx++) { dearly[k][qx] = 0.0; }
      if (t < C::T3 && e < nev)
      {
         const double *dq = qslab + e * (C::NQD * Q3) + Q * row;
#pragma unroll
         for (int k = 0; k < C::NQD; k++) { load_row<Q, true>(dq + k * Q3, dearly[k]); }
      }
   }
   // ---- stage 2 (y): contraction over dy ; task t2 = dx + D*(qz + Q*e) --------------
   if (t < C::T2)
   {
      double a0[D], a1[D];
#pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         a0[dy] = r1[dy * C::SA + t];
         if (DIFF) { a1[dy] = r1[(D + dy) * C::SA + t]; }
      }
      const int dx = t % D, zq = t / D; // zq = qz + Q*e
      double *C0 = r2 + dx * C::SC + Q * zq;
#pragma unroll
      for (int qy = 0; qy < Q; qy++)
      {
         double bb = 0.0, gb = 0.0, bg = 0.0;
         if (COLLOC) { bb = a0[qy < D ? qy : 0]; if (DIFF) { bg = a1[qy < D ? qy : 0]; } }
#pragma unroll
         for (int dy = 0; dy < D; dy++)
         {
            if (!COLLOC)
            {
               bb = fma(bs.B[qy][dy], a0[dy], bb);
               if (DIFF) { bg = fma(bs.B[qy][dy], a1[dy], bg); }
            }
            if (DIFF) { gb = fma(bs.G[qy][dy], a0[dy], gb); }
         }
         C0[qy] = bb;
         if (DIFF)
         {
            C0[D * C::SC + qy] = gb;     // d/dy
            C0[2 * D * C::SC + qy] = bg; // d/dz
         }
      }
   }
   __syncthreads();

   // ---- stage 3 (x, D, x^T) ; task t3 = qy + Q*(qz + Q*e) ---------------------------
   double o0[D], o1[D], o2[D]; // x^T outputs, written after the barrier (buffer reuse)
   if (t < C::T3)
   {
      double c0[D], c1[D], c2[D];
#pragma unroll
      for (int dx = 0; dx < D; dx++)
      {
         c0[dx] = r2[dx * C::SC + t];
         if (DIFF)
         {
            c1[dx] = r2[(D + dx) * C::SC + t];
            c2[dx] = r2[(2 * D + dx) * C::SC + t];
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
      // quadrature data of this x-line -> registers (from the TMA slab or straight from HBM/L2)
      if (use_bar) { mbar_wait(bar, 0); }
      const int e = t / Q2, row = t - e * Q2; // row = qy + Q*qz
      double dv[C::NQD][Q];
      if constexpr (DMODE != 3)
      {
         const double *dq = (C::STAGED ? qd : qslab) + e * (C::NQD * Q3) + Q * row;
         if (C::STAGED || e < nev)
         {
#pragma unroll
            for (int k = 0; k < C::NQD; k++)
            {
               if (DMODE == 0)
               {
#pragma unroll
                  for (int qx = 0; qx < Q; qx++) { dv[k][qx] = dq[k * Q3 + qx]; }
               }
               else { load_row<Q, !C::STAGED>(dq + k * Q3, dv[k]); }
            }
         }
         else
         {
#pragma unroll
            for (int k = 0; k < C::NQD; k++)
#pragma unroll
               for (int qx = 0; qx < Q; qx++) { dv[k][qx] = 0.0; }
         }
      }
      else
      {
#pragma unroll
         for (int k = 0; k < C::NQD; k++)
#pragma unroll
            for (int qx = 0; qx < Q; qx++) { dv[k][qx] = dearly[k][qx]; }
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
      // x^T
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
         o0[dx] = va; o1[dx] = vb; o2[dx] = vc;
      }
   }
   __syncthreads(); // every C read is done: E may overwrite the same region
   if (t < C::T3)
   {
      const int qy = t % Q, zq = t / Q; // zq = qz + Q*e
      double *E0 = r2 + qy * C::SE + D * zq;
#pragma unroll
      for (int dx = 0; dx < D; dx++)
      {
         E0[dx] = o0[dx];
         if (DIFF)
         {
            E0[Q * C::SE + dx] = o1[dx];
            E0[2 * Q * C::SE + dx] = o2[dx];
         }
      }
   }
   __syncthreads();

   // ---- stage 4 (y^T) ; task t2 = dx + D*(qz + Q*e) ---------------------------------
   if (t < C::T2)
   {
      double ea[Q], eb[Q], ec[Q];
#pragma unroll
      for (int qy = 0; qy < Q; qy++)
      {
         ea[qy] = r2[qy * C::SE + t];
         if (DIFF)
         {
            eb[qy] = r2[(Q + qy) * C::SE + t];
            ec[qy] = r2[(2 * Q + qy) * C::SE + t];
         }
      }
      const int dx = t % D, zq = t / D;
      const int e = zq / Q, qz = zq - e * Q;
      double *F0 = r1 + qz * C::SF + dx + D2 * e;
#pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         double p = 0.0, r = 0.0;
         if (COLLOC) { if (DIFF) { p = ea[dy]; r = ec[dy]; } else { p = ea[dy]; } }
#pragma unroll
         for (int qy = 0; qy < Q; qy++)
         {
            if (DIFF)
            {
               p = fma(bs.G[qy][dy], eb[qy], p);
               if (!COLLOC)
               {
                  p = fma(bs.B[qy][dy], ea[qy], p);
                  r = fma(bs.B[qy][dy], ec[qy], r);
               }
            }
            else if (!COLLOC) { p = fma(bs.B[qy][dy], ea[qy], p); }
         }
         F0[D * dy] = p;
         if (DIFF) { F0[Q * C::SF + D * dy] = r; }
      }
   }
   __syncthreads();

   // ---- stage 5 (z^T) + scatter ; task (dx,dy,e) as in stage 1 ----------------------
   {
      const int e = t / D2, r = t - e * D2;
      if ((t < C::T1) && (e < nev))
      {
         double f0[Q], f1[Q];
#pragma unroll
         for (int qz = 0; qz < Q; qz++)
         {
            f0[qz] = r1[qz * C::SF + t];
            if (DIFF) { f1[qz] = r1[(Q + qz) * C::SF + t]; }
         }
         const int64_t base = (e0 + e) * D3 + r;
#pragma unroll
         for (int dz = 0; dz < D; dz++)
         {
            double v = 0.0;
            if (COLLOC) { v = f0[dz]; }
#pragma unroll
            for (int qz = 0; qz < Q; qz++)
            {
               if (!COLLOC) { v = fma(bs.B[qz][dz], f0[qz], v); }
               if (DIFF) { v = fma(bs.G[qz][dz], f1[qz], v); }
            }
            if (FUSED)
            {
               if (gi[dz] >= 0) { atomicAdd(y + gi[dz], v); }
            }
            else { y[base + dz * D2] += v; }
         }
      }
   }
}

// =====================================================================================
// 2D kernel: x -> (y, D, y^T) -> x^T
// =====================================================================================
template <int KIND, int D, int Q, int NEB>
struct Cfg2D
{
   static constexpr bool DIFF = (KIND == B2F_KIND_DIFFUSION);
   static constexpr int NA = DIFF ? 2 : 1;
   static constexpr int NQD = DIFF ? 3 : 1;
   static constexpr int D2 = D * D, Q2 = Q * Q;
   static constexpr int T1 = NEB * D; // x and x^T stage tasks (dy,e)
   static constexpr int T2 = NEB * Q; // y / D / y^T stage tasks (qx,e)
   static constexpr int NT = ((T2 + 31) / 32) * 32;
   static constexpr int SA = pad16(T2, 1); // A[field][dy][t2]
   static constexpr int SF = pad16(T1, 1); // F[field][qx][t1]
   static constexpr int R1 = (NA * D * SA > NA * Q * SF) ? NA * D * SA : NA * Q * SF;
   static constexpr int QD = NEB * NQD * Q2;
   static constexpr size_t SMEM = 16 + sizeof(double) * (size_t)(QD + R1);
   static_assert(T1 <= NT, "one task per thread per stage");
};

template <int KIND, int D, int Q, int NEB, bool FUSED, bool COLLOC>
__global__ void __launch_bounds__(Cfg2D<KIND, D, Q, NEB>::NT)
pa_apply_2d_kernel(const __grid_constant__ Basis<D, Q> bs, const double *__restrict__ qdata,
                   const double *__restrict__ x, double *__restrict__ y,
                   const int *__restrict__ idx, const int64_t ne)
{
   using C = Cfg2D<KIND, D, Q, NEB>;
   constexpr bool DIFF = C::DIFF;
   constexpr int D2 = C::D2, Q2 = C::Q2;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
   double *qd = reinterpret_cast<double *>(smem_raw + 16);
   double *r1 = qd + C::QD;

   const int t = threadIdx.x;
   const int64_t e0 = (int64_t)blockIdx.x * NEB;
   const int nev = (int)((ne - e0) < NEB ? (ne - e0) : NEB);
   const bool use_bar =
      load_qdata<C::NT>(qd, qdata + e0 * (int64_t)(C::NQD * Q2), nev * C::NQD * Q2, bar);

   // stage 1 (x): task (dy,e) owns the x-line: u[dx] -> B u, G u at qx
   int gi[D];
   const int e1 = t / D, dy1 = t - e1 * D;
   {
      double u[D];
#pragma unroll
      for (int dx = 0; dx < D; dx++) { gi[dx] = -1; u[dx] = 0.0; }
      if ((t < C::T1) && (e1 < nev))
      {
         const int64_t base = (e0 + e1) * D2 + dy1 * D;
         if (FUSED)
         {
#pragma unroll
            for (int dx = 0; dx < D; dx++) { gi[dx] = __ldg(idx + base + dx); }
#pragma unroll
            for (int dx = 0; dx < D; dx++) { u[dx] = gi[dx] >= 0 ? __ldg(x + gi[dx]) : 0.0; }
         }
         else
         {
#pragma unroll
            for (int dx = 0; dx < D; dx++) { u[dx] = __ldg(x + base + dx); }
         }
      }
      if (t < C::T1)
      {
         double *A0 = r1 + dy1 * C::SA + Q * e1;
#pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            double vb = 0.0, vg = 0.0;
            if (COLLOC) { vb = u[qx < D ? qx : 0]; }
#pragma unroll
            for (int dx = 0; dx < D; dx++)
            {
               if (!COLLOC) { vb = fma(bs.B[qx][dx], u[dx], vb); }
               if (DIFF) { vg = fma(bs.G[qx][dx], u[dx], vg); }
            }
            A0[qx] = vb;
            if (DIFF) { A0[D * C::SA + qx] = vg; }
         }
      }
   }
   __syncthreads();

   // stage 2 (y, D, y^T): task t2 = qx + Q*e owns the y-line
   double o0[D], o1[D];
   if (t < C::T2)
   {
      double a0[D], a1[D];
#pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         a0[dy] = r1[dy * C::SA + t];
         if (DIFF) { a1[dy] = r1[(D + dy) * C::SA + t]; }
      }
      double gx[Q], gy[Q];
#pragma unroll
      for (int qy = 0; qy < Q; qy++)
      {
         double vx = 0.0, vy = 0.0;
         if (COLLOC) { if (DIFF) { vx = a1[qy < D ? qy : 0]; } else { vx = a0[qy < D ? qy : 0]; } }
#pragma unroll
         for (int dy = 0; dy < D; dy++)
         {
            if (DIFF)
            {
               if (!COLLOC) { vx = fma(bs.B[qy][dy], a1[dy], vx); } // d/dx: G_x B_y
               vy = fma(bs.G[qy][dy], a0[dy], vy);                  // d/dy: B_x G_y
            }
            else if (!COLLOC) { vx = fma(bs.B[qy][dy], a0[dy], vx); }
         }
         gx[qy] = vx; gy[qy] = vy;
      }
      if (use_bar) { mbar_wait(bar, 0); }
      const int e = t / Q, qx = t - e * Q;
      const double *dq = qd + e * (C::NQD * Q2) + qx;
      if (DIFF)
      {
#pragma unroll
         for (int qy = 0; qy < Q; qy++)
         {
            const double O11 = dq[Q * qy], O12 = dq[Q2 + Q * qy], O22 = dq[2 * Q2 + Q * qy];
            const double ux = gx[qy], uy = gy[qy];
            gx[qy] = (O11 * ux) + (O12 * uy);
            gy[qy] = (O12 * ux) + (O22 * uy);
         }
      }
      else
      {
#pragma unroll
         for (int qy = 0; qy < Q; qy++) { gx[qy] *= dq[Q * qy]; }
      }
#pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         double va = 0.0, vb = 0.0;
         if (COLLOC) { va = gx[dy]; }
#pragma unroll
         for (int qy = 0; qy < Q; qy++)
         {
            if (!COLLOC) { va = fma(bs.B[qy][dy], gx[qy], va); }
            if (DIFF) { vb = fma(bs.G[qy][dy], gy[qy], vb); }
         }
         o0[dy] = va; o1[dy] = vb; // o0 -> G_x^T, o1 -> B_x^T   (diffusion)
      }
   }
   __syncthreads();
   if (t < C::T2)
   {
      const int e = t / Q, qx = t - e * Q;
      double *F0 = r1 + qx * C::SF + D * e;
#pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         F0[dy] = o0[dy];
         if (DIFF) { F0[Q * C::SF + dy] = o1[dy]; }
      }
   }
   __syncthreads();

   // stage 3 (x^T) + scatter: task (dy,e)
   if ((t < C::T1) && (e1 < nev))
   {
      double f0[Q], f1[Q];
#pragma unroll
      for (int qx = 0; qx < Q; qx++)
      {
         f0[qx] = r1[qx * C::SF + t];
         if (DIFF) { f1[qx] = r1[(Q + qx) * C::SF + t]; }
      }
      const int64_t base = (e0 + e1) * D2 + dy1 * D;
#pragma unroll
      for (int dx = 0; dx < D; dx++)
      {
         double v = 0.0;
         if (COLLOC && !DIFF) { v = f0[dx]; }
         if (COLLOC && DIFF) { v = f1[dx]; }
#pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            if (DIFF)
            {
               v = fma(bs.G[qx][dx], f0[qx], v);
               if (!COLLOC) { v = fma(bs.B[qx][dx], f1[qx], v); }
            }
            else if (!COLLOC) { v = fma(bs.B[qx][dx], f0[qx], v); }
         }
         if (FUSED)
         {
            if (gi[dx] >= 0) { atomicAdd(y + gi[dx], v); }
         }
         else { y[base + dx] += v; }
      }
   }
}

// ---- host-side launch helpers --------------------------------------------------------
template <int D, int Q>
inline void fill_basis(const b2f_pa *pa, Basis<D, Q> &bs)
{
   for (int q = 0; q < Q; q++)
      for (int d = 0; d < D; d++)
      {
         bs.B[q][d] = pa->B[q + Q * d];
         bs.G[q][d] = pa->G[q + Q * d];
      }
}

// FUSEDSEL: 0 = E-vector form only, 1 = fused L-vector form only, 2 = both (runtime idx)
template <int KIND, int D, int Q, int NEB, int DMODE = 0, int MINB = 1, int FUSEDSEL = 2>
int launch_3d(b2f_pa *pa, const double *x, double *y, const int *idx)
{
   using C = Cfg3D<KIND, D, Q, NEB, DMODE>;
   b2f_ctx *ctx = pa->ctx;
   if ((int)C::SMEM > ctx->max_smem_optin)
   {
      set_error("pa_apply_3d<D=%d,Q=%d,NEB=%d> needs %zu B of shared memory", D, Q, NEB, C::SMEM);
      return B2F_ERR_UNSUPPORTED;
   }
   Basis<D, Q> bs;
   fill_basis(pa, bs);
   const int64_t grid = cdiv(pa->ne, NEB);
   auto go = [&](auto kern) -> int
   {
      // once per (device, instantiation): opt in to the shared memory size and find how many CTAs
      // are resident on the device -- that is the L2 prefetch distance (the slab a CTA prefetches
      // belongs to the CTA that takes over its slot)
      int resident = 0;
      if (int rc = kernel_residency(ctx, kern, C::NT, C::SMEM, resident)) { return rc; }
      const char *pfs = getenv("B2F_PF");
      const int pf = pfs ? atoi(pfs) : resident;
      timing_mark(ctx);
      kern<<<(unsigned)grid, C::NT, C::SMEM, ctx->stream>>>(bs, pa->data, x, y, idx, pa->ne, pf);
      timing_mark(ctx);
      B2F_LAUNCH_CHECK(ctx);
      return B2F_OK;
   };
   const bool fused = idx != nullptr;
   if ((fused && FUSEDSEL == 0) || (!fused && FUSEDSEL == 1))
   {
      set_error("pa_apply_3d: kernel form not instantiated");
      return B2F_ERR_UNSUPPORTED;
   }
   constexpr bool CL = (D == Q);
   if (pa->collocated && CL)
   {
      if constexpr (FUSEDSEL != 0) { if (fused) { return go(pa_apply_3d_kernel<KIND, D, Q, NEB, true, CL, DMODE, MINB>); } }
      if constexpr (FUSEDSEL != 1) { if (!fused) { return go(pa_apply_3d_kernel<KIND, D, Q, NEB, false, CL, DMODE, MINB>); } }
   }
   if constexpr (FUSEDSEL != 0) { if (fused) { return go(pa_apply_3d_kernel<KIND, D, Q, NEB, true, false, DMODE, MINB>); } }
   if constexpr (FUSEDSEL != 1) { if (!fused) { return go(pa_apply_3d_kernel<KIND, D, Q, NEB, false, false, DMODE, MINB>); } }
   return B2F_ERR_UNSUPPORTED;
}

template <int KIND, int D, int Q, int NEB>
int launch_2d(b2f_pa *pa, const double *x, double *y, const int *idx)
{
   using C = Cfg2D<KIND, D, Q, NEB>;
   b2f_ctx *ctx = pa->ctx;
   Basis<D, Q> bs;
   fill_basis(pa, bs);
   const int64_t grid = cdiv(pa->ne, NEB);
   auto go = [&](auto kern) -> int
   {
      B2F_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
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
      return fused ? go(pa_apply_2d_kernel<KIND, D, Q, NEB, true, CL>)
             : go(pa_apply_2d_kernel<KIND, D, Q, NEB, false, CL>);
   }
   return fused ? go(pa_apply_2d_kernel<KIND, D, Q, NEB, true, false>)
          : go(pa_apply_2d_kernel<KIND, D, Q, NEB, false, false>);
}

// per-(D,Q) entry points, defined in pa_apply_{3d_diff,3d_mass,2d}.cu
int pa_apply_3d_diffusion(b2f_pa *pa, const double *x, double *y, const int *idx);
int pa_apply_3d_mass(b2f_pa *pa, const double *x, double *y, const int *idx);
int pa_apply_2d(b2f_pa *pa, const double *x, double *y, const int *idx);

} // namespace b2f
