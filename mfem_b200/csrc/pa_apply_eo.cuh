// mfem_b200/csrc/pa_apply_eo.cuh -- 3D PA operator kernel, even-odd sum factorisation.
//
// Same operator as pa_apply.cuh (MFEM's SmemPADiffusionApply3D / SmemPAMassApply3D,
// fem/integ/bilininteg_diffusion_kernels.hpp:989-1215, bilininteg_mass_kernels.hpp:815-1165,
// fused with ElementRestriction::Mult / MultTranspose, fem/restriction.cpp:116,159), with two
// changes that cut what actually limits the kernel on B200 (profiles/: the LSU / shared-memory
// pipe and instruction issue at the 1 kW power cap, not HBM):
//
//  * EVEN-ODD decomposition of every 1D contraction.  The H1 basis on symmetric nodes and a
//    symmetric quadrature rule satisfies B[Q-1-q][D-1-d] = B[q][d], G[Q-1-q][D-1-d] =
//    -G[q][d].  Splitting a line into its symmetric and antisymmetric halves
//    (s_i = u_i + u_{D-1-i}, a_i = u_i - u_{D-1-i}) halves the multiply-adds and -- as every
//    DFMA needs its own basis entry from the constant bank -- halves the constant loads:
//    per 1D transform D*Q FMAs + D*Q constant loads become D*Q/2 FMAs + D*Q/2 loads + (D+Q)
//    adds.  Results differ from the plain contraction by rounding only (~1e-16).
//  * the shared-memory layout of pa_eo_cfg.cuh: stage order y -> z -> x with
//    yq = qy + Q*e the fastest lane AND address index of stages 2-4 (consecutive lanes touch
//    consecutive words: no bank conflicts, no transposition), the x-stage overwrites its
//    own input column in place (no staging registers, one barrier fewer), quadrature-data
//    rows are read with 128-bit LDS from per-element TMA slabs padded so that lanes of
//    neighbouring elements hit different bank groups (Q = 8: swizzled tensor copies instead).
//
// Streams: one TMA bulk copy per element (cp.async.bulk + mbarrier complete_tx) at kernel
// entry; an L2 bulk prefetch of the slab and gather indices of the CTA that will take over
// this SM slot (`pf` blocks ahead).  Requires the symmetry above (checked on the host when
// the integrator is set up; otherwise pa_apply.cuh's general kernel runs).
#pragma once
#include <map>
#include <cuda.h>
#include <mutex>
#include "pa_eo_cfg.cuh"

namespace b2f
{

template <int D, int Q>
struct EOBasis
{
   static constexpr int DH = (D + 1) / 2, QH = (Q + 1) / 2;
   // M[q][i] = S[q][i] (s-part) and A[q][i] (a-part) for q < QH, i < DH; the middle column
   // (D odd) is stored in S, its A entry is 0
   double Bs[QH][DH], Ba[QH][DH], Gs[QH][DH], Ga[QH][DH];
};

// s_i = u_i + sg*u_{N-1-i}, a_i = u_i - sg*u_{N-1-i}; middle (N odd): s = a = u
template <int N, bool NEG>
__device__ __forceinline__ void eo_split(const double (&u)[N], double (&s)[(N + 1) / 2],
                                         double (&a)[(N + 1) / 2])
{
#pragma unroll
   for (int i = 0; i < N / 2; i++)
   {
      s[i] = NEG ? u[i] - u[N - 1 - i] : u[i] + u[N - 1 - i];
      a[i] = NEG ? u[i] + u[N - 1 - i] : u[i] - u[N - 1 - i];
   }
   if (N % 2) { s[N / 2] = u[N / 2]; a[N / 2] = u[N / 2]; }
}

// forward D -> Q:  es[q] += sum_i S[q][i] s_i ; ea[q] += sum_{i<D/2} A[q][i] a_i
// out[q] = es + ea ; out[Q-1-q] = sg (es - ea).  ANTI: the matrix is antisymmetric (G).
template <int D, int Q, bool ANTI>
__device__ __forceinline__ void eo_fwd(const double (&S)[(Q + 1) / 2][(D + 1) / 2],
                                       const double (&A)[(Q + 1) / 2][(D + 1) / 2],
                                       const double (&s)[(D + 1) / 2], const double (&a)[(D + 1) / 2],
                                       double (&out)[Q])
{
   constexpr int DH = (D + 1) / 2, QH = (Q + 1) / 2, dh = D / 2, qh = Q / 2;
#pragma unroll
   for (int q = 0; q < QH; q++)
   {
      double es = 0.0, ea = 0.0;
      const bool mid = (Q % 2) && q == qh; // middle row: sym -> A row is 0, anti -> S row is 0
      if (!(mid && ANTI))
      {
#pragma unroll
         for (int i = 0; i < DH; i++) { es = fma(S[q][i], s[i], es); }
      }
      if (!(mid && !ANTI))
      {
#pragma unroll
         for (int i = 0; i < dh; i++) { ea = fma(A[q][i], a[i], ea); }
      }
      out[q] = es + ea;
      if (!mid) { out[Q - 1 - q] = ANTI ? ea - es : es - ea; }
   }
}

// backward Q -> D, accumulating in the split space:
// es[i] += sum_q S[q][i] s_q ; ea[i] += sum_q A[q][i] a_q   (s,a split with sign sg)
template <int D, int Q, bool ANTI>
__device__ __forceinline__ void eo_bwd_acc(const double (&S)[(Q + 1) / 2][(D + 1) / 2],
                                           const double (&A)[(Q + 1) / 2][(D + 1) / 2],
                                           const double (&s)[(Q + 1) / 2], const double (&a)[(Q + 1) / 2],
                                           double (&es)[(D + 1) / 2], double (&ea)[(D + 1) / 2])
{
   constexpr int DH = (D + 1) / 2, QH = (Q + 1) / 2, dh = D / 2, qh = Q / 2;
   // middle row (Q odd): sym -> contributes to es only, anti -> to ea only
   constexpr int QS = ANTI ? qh : QH, QA = ANTI ? QH : qh;
#pragma unroll
   for (int i = 0; i < DH; i++)
   {
#pragma unroll
      for (int q = 0; q < QS; q++) { es[i] = fma(S[q][i], s[q], es[i]); }
      if (i < dh)
      {
#pragma unroll
         for (int q = 0; q < QA; q++) { ea[i] = fma(A[q][i], a[q], ea[i]); }
      }
   }
}

template <int D>
__device__ __forceinline__ void eo_bwd_out(const double (&es)[(D + 1) / 2],
                                           const double (&ea)[(D + 1) / 2], double (&out)[D])
{
#pragma unroll
   for (int i = 0; i < D / 2; i++)
   {
      out[i] = es[i] + ea[i];
      out[D - 1 - i] = es[i] - ea[i];
   }
   if (D % 2) { out[D / 2] = es[D / 2]; }
}

template <int N>
__device__ __forceinline__ void zero(double (&v)[N])
{
#pragma unroll
   for (int i = 0; i < N; i++) { v[i] = 0.0; }
}

// ---- the five stages as device functions (shared by the one-shot and persistent kernels) ----
template <int KIND, int D, int Q, int NEB, bool COLLOC, bool WITH_ENERGY = false>
struct EOStages
{
   using C = PipeCfg<KIND, D, Q, NEB>;
   static constexpr bool DIFF = C::DIFF;
   static constexpr int Q3 = C::Q3, YQ = C::YQ;
   static constexpr int DH = (D + 1) / 2, QH = (Q + 1) / 2;
   using BS = EOBasis<D, Q>;

   // stage 1 (y): line (dx, dz, e): u[dy] -> B u, G u at qy, written to A at a1 + qy
   static __device__ __forceinline__ void stage1(const BS &bs, const double (&u)[D], double *A0)
   {
      double s[DH], a[DH], vb[Q], vg[Q];
      eo_split<D, false>(u, s, a);
      if (COLLOC)
      {
#pragma unroll
         for (int q = 0; q < Q; q++) { vb[q] = u[q < D ? q : 0]; }
      }
      else { eo_fwd<D, Q, false>(bs.Bs, bs.Ba, s, a, vb); }
      if (DIFF) { eo_fwd<D, Q, true>(bs.Gs, bs.Ga, s, a, vg); }
      if constexpr (C::VEC1 && C::TMA_OK)
      {
         double2 *A2 = reinterpret_cast<double2 *>(A0);
#pragma unroll
         for (int j = 0; j < Q / 2; j++)
         {
            A2[j] = make_double2(vb[2 * j], vb[2 * j + 1]);
            if (DIFF) { A2[(D * C::SA) / 2 + j] = make_double2(vg[2 * j], vg[2 * j + 1]); }
         }
      }
      else
      {
#pragma unroll
         for (int qy = 0; qy < Q; qy++)
         {
            A0[qy] = vb[qy];
            if (DIFF) { A0[D * C::SA + qy] = vg[qy]; }
         }
      }
   }

   // stage 2 (z): line t2 = (yq, dx), contraction over dz
   static __device__ __forceinline__ void stage2(const BS &bs, const double *r1, double *r2, int t)
   {
      double a0[D], a1v[D];
#pragma unroll
      for (int dz = 0; dz < D; dz++)
      {
         a0[dz] = r1[t + C::SA * dz];
         if (DIFF) { a1v[dz] = r1[t + C::SA * (D + dz)]; }
      }
      double s0[DH], d0[DH], bb[Q], gyv[Q], gzv[Q];
      eo_split<D, false>(a0, s0, d0);
      if (COLLOC)
      {
#pragma unroll
         for (int q = 0; q < Q; q++) { bb[q] = a0[q < D ? q : 0]; if (DIFF) { gyv[q] = a1v[q < D ? q : 0]; } }
      }
      else { eo_fwd<D, Q, false>(bs.Bs, bs.Ba, s0, d0, bb); }
      if (DIFF)
      {
         eo_fwd<D, Q, true>(bs.Gs, bs.Ga, s0, d0, gzv);
         if (!COLLOC)
         {
            double s1[DH], d1[DH];
            eo_split<D, false>(a1v, s1, d1);
            eo_fwd<D, Q, false>(bs.Bs, bs.Ba, s1, d1, gyv);
         }
      }
      const int yq = t % YQ, dx = t / YQ;
      double *C0 = r2 + yq + C::SC * dx;
#pragma unroll
      for (int qz = 0; qz < Q; qz++)
      {
         C0[YQ * qz] = bb[qz];
         if (DIFF)
         {
            C0[D * C::SC + YQ * qz] = gyv[qz];     // d/dy
            C0[2 * D * C::SC + YQ * qz] = gzv[qz]; // d/dz
         }
      }
   }

   // stage 3 (x, D, x^T) in place: line t3 = (yq, qz).  Returns this line's share of the
   // energy x.(Ax) = sum_q (grad u)^T D (grad u)  (mass: sum_q u D u), free at this point
   static __device__ __forceinline__ double stage3(const BS &bs, double *r2, const double *qd, int t,
                                                   uint64_t *bar, uint32_t parity)
   {
      double energy = 0.0;
      const int yq = t % YQ, qz = t / YQ;
      const int e = yq / Q, qy = yq - e * Q;
      double *C0 = r2 + yq + YQ * qz;
      double gx[Q], gy[Q], gz[Q];
      {
         double c0[D], s[DH], a[DH];
#pragma unroll
         for (int dx = 0; dx < D; dx++) { c0[dx] = C0[C::SC * dx]; }
         eo_split<D, false>(c0, s, a);
         if (DIFF) { eo_fwd<D, Q, true>(bs.Gs, bs.Ga, s, a, gx); }
         else if (COLLOC)
         {
#pragma unroll
            for (int q = 0; q < Q; q++) { gx[q] = c0[q < D ? q : 0]; }
         }
         else { eo_fwd<D, Q, false>(bs.Bs, bs.Ba, s, a, gx); }
      }
      if (DIFF)
      {
         double c1[D], c2[D];
#pragma unroll
         for (int dx = 0; dx < D; dx++)
         {
            c1[dx] = C0[C::SC * (D + dx)];
            c2[dx] = C0[C::SC * (2 * D + dx)];
         }
         if (COLLOC)
         {
#pragma unroll
            for (int q = 0; q < Q; q++) { gy[q] = c1[q < D ? q : 0]; gz[q] = c2[q < D ? q : 0]; }
         }
         else
         {
            double s[DH], a[DH];
            eo_split<D, false>(c1, s, a);
            eo_fwd<D, Q, false>(bs.Bs, bs.Ba, s, a, gy);
            eo_split<D, false>(c2, s, a);
            eo_fwd<D, Q, false>(bs.Bs, bs.Ba, s, a, gz);
         }
      }
      // quadrature data of this line: NQD rows of Q doubles from the TMA-filled slab
      if constexpr (C::TMA_OK) { mbar_wait(bar, C::ALIAS ? 0u : parity); }
      if constexpr (C::ALIAS) { mbar_wait(bar + 1, parity); } // the slab's tail (pa_eo_cfg.cuh: ALIAS)
      const double *dq = qd + e * C::DS + Q * (qy + Q * qz);
      if (DIFF)
      {
         double dv[C::NQD][Q];
#pragma unroll
         for (int k = 0; k < C::NQD; k++)
         {
            if constexpr (C::SWZ) { load_row_swz<Q>(dq + k * Q3, swz_phase<Q>(e, qy, qz), dv[k]); }
            else if constexpr (C::VEC) { load_row<Q, false>(dq + k * Q3, dv[k]); }
            else
            {
#pragma unroll
               for (int qx = 0; qx < Q; qx++) { dv[k][qx] = dq[k * Q3 + qx]; }
            }
         }
#pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            const double O11 = dv[0][qx], O12 = dv[1 % C::NQD][qx], O13 = dv[2 % C::NQD][qx];
            const double O22 = dv[3 % C::NQD][qx], O23 = dv[4 % C::NQD][qx], O33 = dv[5 % C::NQD][qx];
            const double ux = gx[qx], uy = gy[qx], uz = gz[qx];
            gx[qx] = (O11 * ux) + (O12 * uy) + (O13 * uz);
            gy[qx] = (O12 * ux) + (O22 * uy) + (O23 * uz);
            gz[qx] = (O13 * ux) + (O23 * uy) + (O33 * uz);
            if (WITH_ENERGY) { energy = fma(ux, gx[qx], fma(uy, gy[qx], fma(uz, gz[qx], energy))); }
         }
      }
      else
      {
         double dv[Q];
         if constexpr (C::SWZ) { load_row_swz<Q>(dq, swz_phase<Q>(e, qy, qz), dv); }
         else if constexpr (C::VEC) { load_row<Q, false>(dq, dv); }
         else
         {
#pragma unroll
            for (int qx = 0; qx < Q; qx++) { dv[qx] = dq[qx]; }
         }
#pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            const double ux = gx[qx];
            gx[qx] *= dv[qx];
            if (WITH_ENERGY) { energy = fma(ux, gx[qx], energy); }
         }
      }
      // x^T, written over this thread's own input column
      {
         double s[QH], a[QH], es[DH], ea[DH], o[D];
         zero(es); zero(ea);
         if (DIFF)
         {
            eo_split<Q, true>(gx, s, a);
            eo_bwd_acc<D, Q, true>(bs.Gs, bs.Ga, s, a, es, ea);
            eo_bwd_out<D>(es, ea, o);
         }
         else if (COLLOC)
         {
#pragma unroll
            for (int dx = 0; dx < D; dx++) { o[dx] = gx[dx]; }
         }
         else
         {
            eo_split<Q, false>(gx, s, a);
            eo_bwd_acc<D, Q, false>(bs.Bs, bs.Ba, s, a, es, ea);
            eo_bwd_out<D>(es, ea, o);
         }
#pragma unroll
         for (int dx = 0; dx < D; dx++) { C0[C::SC * dx] = o[dx]; }
      }
      if (DIFF)
      {
         double s[QH], a[QH], es[DH], ea[DH], o[D];
         if (COLLOC)
         {
#pragma unroll
            for (int dx = 0; dx < D; dx++) { C0[C::SC * (D + dx)] = gy[dx]; C0[C::SC * (2 * D + dx)] = gz[dx]; }
         }
         else
         {
            zero(es); zero(ea);
            eo_split<Q, false>(gy, s, a);
            eo_bwd_acc<D, Q, false>(bs.Bs, bs.Ba, s, a, es, ea);
            eo_bwd_out<D>(es, ea, o);
#pragma unroll
            for (int dx = 0; dx < D; dx++) { C0[C::SC * (D + dx)] = o[dx]; }
            zero(es); zero(ea);
            eo_split<Q, false>(gz, s, a);
            eo_bwd_acc<D, Q, false>(bs.Bs, bs.Ba, s, a, es, ea);
            eo_bwd_out<D>(es, ea, o);
#pragma unroll
            for (int dx = 0; dx < D; dx++) { C0[C::SC * (2 * D + dx)] = o[dx]; }
         }
      }
      return energy;
   }

   // stage 4 (z^T): line t2 = (yq, dx)
   static __device__ __forceinline__ void stage4(const BS &bs, const double *r2, double *r1, int t)
   {
      const int yq = t % YQ, dx = t / YQ;
      const double *E0 = r2 + yq + C::SC * dx;
      double ev[Q], s[QH], a[QH], es[DH], ea[DH], p0[D], p1[D];
      zero(es); zero(ea);
#pragma unroll
      for (int qz = 0; qz < Q; qz++) { ev[qz] = E0[YQ * qz]; }
      if (COLLOC)
      {
#pragma unroll
         for (int dz = 0; dz < D; dz++) { p0[dz] = ev[dz]; }
      }
      else
      {
         eo_split<Q, false>(ev, s, a);
         eo_bwd_acc<D, Q, false>(bs.Bs, bs.Ba, s, a, es, ea);
      }
      if (DIFF)
      {
#pragma unroll
         for (int qz = 0; qz < Q; qz++) { ev[qz] = E0[2 * D * C::SC + YQ * qz]; }
         eo_split<Q, true>(ev, s, a);
         eo_bwd_acc<D, Q, true>(bs.Gs, bs.Ga, s, a, es, ea); // + G^T (d/dz part)
      }
      if (COLLOC)
      {
         if (DIFF)
         {
            double g[D];
            eo_bwd_out<D>(es, ea, g);
#pragma unroll
            for (int dz = 0; dz < D; dz++) { p0[dz] += g[dz]; }
         }
      }
      else { eo_bwd_out<D>(es, ea, p0); }
      if (DIFF)
      {
#pragma unroll
         for (int qz = 0; qz < Q; qz++) { ev[qz] = E0[D * C::SC + YQ * qz]; }
         if (COLLOC)
         {
#pragma unroll
            for (int dz = 0; dz < D; dz++) { p1[dz] = ev[dz]; }
         }
         else
         {
            zero(es); zero(ea);
            eo_split<Q, false>(ev, s, a);
            eo_bwd_acc<D, Q, false>(bs.Bs, bs.Ba, s, a, es, ea);
            eo_bwd_out<D>(es, ea, p1);
         }
      }
#pragma unroll
      for (int dz = 0; dz < D; dz++)
      {
         r1[t + C::SA * dz] = p0[dz];
         if (DIFF) { r1[t + C::SA * (D + dz)] = p1[dz]; }
      }
   }

   // stage 5 (y^T): line (dx, dz, e): F at F0 + qy -> v[dy]
   static __device__ __forceinline__ void stage5(const BS &bs, const double *F0, double (&v)[D])
   {
      double f[Q], s[QH], a[QH], es[DH], ea[DH];
      zero(es); zero(ea);
      constexpr bool V1 = C::VEC1 && C::TMA_OK;
      if constexpr (V1)
      {
         const double2 *F2 = reinterpret_cast<const double2 *>(F0);
#pragma unroll
         for (int j = 0; j < Q / 2; j++) { const double2 v2 = F2[j]; f[2 * j] = v2.x; f[2 * j + 1] = v2.y; }
      }
      else
      {
#pragma unroll
         for (int qy = 0; qy < Q; qy++) { f[qy] = F0[qy]; }
      }
      if (COLLOC)
      {
#pragma unroll
         for (int dy = 0; dy < D; dy++) { v[dy] = f[dy]; }
      }
      else
      {
         eo_split<Q, false>(f, s, a);
         eo_bwd_acc<D, Q, false>(bs.Bs, bs.Ba, s, a, es, ea);
      }
      if (DIFF)
      {
         if constexpr (V1)
         {
            const double2 *F2 = reinterpret_cast<const double2 *>(F0 + D * C::SA);
#pragma unroll
            for (int j = 0; j < Q / 2; j++) { const double2 v2 = F2[j]; f[2 * j] = v2.x; f[2 * j + 1] = v2.y; }
         }
         else
         {
#pragma unroll
            for (int qy = 0; qy < Q; qy++) { f[qy] = F0[D * C::SA + qy]; }
         }
         eo_split<Q, true>(f, s, a);
         eo_bwd_acc<D, Q, true>(bs.Gs, bs.Ga, s, a, es, ea);
      }
      if (COLLOC)
      {
         if (DIFF)
         {
            double g[D];
            eo_bwd_out<D>(es, ea, g);
#pragma unroll
            for (int dy = 0; dy < D; dy++) { v[dy] += g[dy]; }
         }
      }
      else { eo_bwd_out<D>(es, ea, v); }
   }
};

// stage 1/5 lane -> (e, dx, dz)
template <int D, int NEB, int ORDER>
__device__ __forceinline__ void eo_line_coords(int t, int &e1, int &dx1, int &dz1)
{
   if (ORDER == 0) { e1 = t % NEB; dx1 = (t / NEB) % D; dz1 = t / (NEB * D); }
   else if (ORDER == 1) { dx1 = t % D; dz1 = (t / D) % D; e1 = t / (D * D); }
   else if (ORDER == 2) { e1 = t % NEB; dz1 = (t / NEB) % D; dx1 = t / (NEB * D); }
   else { dz1 = t % D; e1 = (t / D) % NEB; dx1 = t / (NEB * D); }
}

// ---- one tile per CTA; next CTA's slab prefetched into L2 --------------------------------------
// DOT: additionally accumulate x . (A x) into dot_acc[64] -- the PCG's (d, A d) without another
// pass over the vectors (IterativeSolver::Dot, linalg/solvers.cpp:58).  It is evaluated as the
// energy sum_q (grad x)^T D (grad x) at the quadrature points, where both factors are in
// registers anyway (3 FMAs per point); absent elements of a ragged tile contribute nothing.
// vector spaces (vdim > 1: CEED BP2/4/6, Vector{Mass,Diffusion}Integrator): the components are
// processed one after the other while the element's quadrature data stays in shared memory
struct VecArgs
{
   int vd;               // number of components
   long long xs, xo;     // component c of L-dof g lives at x[g*xs + c*xo]
   long long ics;        // gather index of component c starts at idx + c*ics (per-component masks)
};

template <int KIND, int D, int Q, int NEB, bool FUSED, bool COLLOC, int MINB, bool DOT = false,
          bool VEC = false>
__global__ void __launch_bounds__(PipeCfg<KIND, D, Q, NEB>::NT, MINB)
pa_apply_3d_eo_kernel(const __grid_constant__ EOBasis<D, Q> bs, const double *__restrict__ qdata,
                      const double *__restrict__ x, double *__restrict__ y,
                      const int *__restrict__ idx, const int64_t ne, const int pf,
                      double *__restrict__ dot_acc, const int64_t e_first, const VecArgs va,
                      const __grid_constant__ TmapArg<PipeCfg<KIND, D, Q, NEB>::SWZ> tmap)
{
   using C = PipeCfg<KIND, D, Q, NEB>;
   using S = EOStages<KIND, D, Q, NEB, COLLOC, DOT>;
   constexpr int D2 = C::D2, D3 = C::D3, Q3 = C::Q3, YQ = C::YQ;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   // swizzled slabs: the quadrature data leads, at a 1024-byte aligned address (the swizzle is a
   // function of the shared-memory address bits); the mbarrier follows the buffers
   unsigned char *sbase = C::SWZ ? smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) : smem_raw;
   double *qd = reinterpret_cast<double *>(C::SWZ ? sbase : sbase + 16);
   double *r1 = C::ALIAS ? qd + (C::NQD * Q3 - C::QB) : qd + C::QD; // ALIAS: over the slab's tail
   double *r2 = C::ALIAS ? qd + C::QD : r1 + C::R1;
   uint64_t *bar = reinterpret_cast<uint64_t *>(C::SWZ ? reinterpret_cast<unsigned char *>(r2 + C::R2) : sbase);

   const int t = threadIdx.x;
   const int64_t e0 = e_first + (int64_t)blockIdx.x * NEB; // elements [e_first, ne) of this launch
   const int nev = (int)((ne - e0) < NEB ? (ne - e0) : NEB);

   int e1, dx1, dz1;
   eo_line_coords<D, NEB, C::L.order>(t, e1, dx1, dz1);
   const bool has1 = t < C::T1;
   const bool act1 = has1 && e1 < nev;
   const int loc1 = dx1 + D2 * dz1;                // + D*dy : position inside the element
   const int a1 = Q * e1 + YQ * dx1 + C::SA * dz1; // + qy   : A/F address of this line

   const int ncomp = VEC ? va.vd : 1;
   for (int c = 0; c < ncomp; c++)
   {
      int gi[D];
      {
         double u[D];
#pragma unroll
         for (int dy = 0; dy < D; dy++) { gi[dy] = -1; u[dy] = 0.0; }
         if constexpr (FUSED)
         {
            // the gather indices head the longest dependent chain (index -> x -> stage 1): their
            // loads are issued before anything else, the copy set-up below runs in their shadow
            if (act1)
            {
               const int *ip = idx + (VEC ? c * va.ics : 0) + (e0 + e1) * D3 + loc1;
#pragma unroll
               for (int dy = 0; dy < D; dy++) { gi[dy] = __ldg(ip + D * dy); }
            }
         }
         if (c == 0)
         {
         // quadrature-data stream: TMA per element + L2 prefetch for the CTA `pf` blocks on
         const double *qslab = qdata + e0 * (int64_t)(C::NQD * Q3);
         if constexpr (C::TMA_OK)
         {
            if (t == 0)
            {
               mbar_init(bar, 1);
               if constexpr (C::ALIAS)
               {
                  // head of the slab now, tail after the z-stage (second barrier)
                  mbar_init(bar + 1, 1);
                  mbar_expect_tx(bar, (uint32_t)(C::NQD * Q3 - C::QB) * 8u);
                  bulk_g2s(qd, qslab, (uint32_t)(C::NQD * Q3 - C::QB) * 8u, bar);
               }
               else
               {
               mbar_expect_tx(bar, (uint32_t)(nev * C::NQD * Q3) * 8u);
               for (int e = 0; e < nev; e++)
               {
                  if constexpr (C::SWZ)
                  {
#pragma unroll
                     for (int b = 0; b < C::NBOX; b++)
                     {
                        tensor_g2s_3d(qd + e * C::DS + b * (C::BOX_ROWS * Q), &tmap, 0, b * C::BOX_ROWS,
                                      (int)(e0 + e), bar);
                     }
                  }
                  else
                  {
                     bulk_g2s(qd + e * C::DS, qslab + (int64_t)e * (C::NQD * Q3), (uint32_t)(C::NQD * Q3) * 8u, bar);
                  }
               }
               }
            }
         }
         else
         {
            for (int i = t; i < nev * C::NQD * Q3; i += C::NT) { cp_async8(qd + i, qslab + i); }
         }
         if (pf > 0 && t == 32 % C::NT)
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
         if (act1)
         {
            if constexpr (FUSED)
            {
#pragma unroll
               for (int dy = 0; dy < D; dy++)
               {
                  const long long a = VEC ? gi[dy] * va.xs + c * va.xo : gi[dy];
                  u[dy] = gi[dy] >= 0 ? __ldg(x + a) : 0.0;
               }
            }
            else
            {
               const double *xe = x + ((e0 + e1) * ncomp + c) * D3 + loc1;
#pragma unroll
               for (int dy = 0; dy < D; dy++) { u[dy] = __ldg(xe + D * dy); }
            }
         }
         if (has1) { S::stage1(bs, u, r1 + a1); }
      }
      __syncthreads();
      if (t < C::T2) { S::stage2(bs, r1, r2, t); }
      if constexpr (!C::TMA_OK) { cp_async_wait_all(); }
      __syncthreads();
      if constexpr (C::ALIAS)
      {
         // the y-stage buffer is dead: bring in the slab's tail over it (once per component: the
         // z^T-stage overwrites it again)
         if (t == 0)
         {
            fence_proxy_async(); // the region was just accessed through the generic proxy
            mbar_expect_tx(bar + 1, (uint32_t)C::QB * 8u);
            bulk_g2s(r1, qdata + e0 * (int64_t)(C::NQD * Q3) + (C::NQD * Q3 - C::QB), (uint32_t)C::QB * 8u, bar + 1);
         }
      }
      {
         double part = 0.0;
         if (t < C::T3)
         {
            part = S::stage3(bs, r2, qd, t, bar, C::ALIAS ? (uint32_t)(c & 1) : 0u);
            if ((t % C::YQ) / Q >= nev) { part = 0.0; } // lines of a ragged tile: absent elements
         }
         if constexpr (DOT)
         {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { part += __shfl_xor_sync(0xffffffffu, part, o); }
            if ((t & 31) == 0) { atomicAdd(dot_acc + ((blockIdx.x + (t >> 5)) & 63), part); }
         }
      }
      __syncthreads();
      if (t < C::T2) { S::stage4(bs, r2, r1, t); }
      __syncthreads();
      if (act1)
      {
         double v[D];
         S::stage5(bs, r1 + a1, v);
         double *ye = y + ((e0 + e1) * ncomp + c) * D3 + loc1; // E-vector form
#pragma unroll
         for (int dy = 0; dy < D; dy++)
         {
            if (FUSED)
            {
               const long long a = VEC ? gi[dy] * va.xs + c * va.xo : gi[dy];
               if (gi[dy] >= 0) { atomicAdd(y + a, v[dy]); }
            }
            else { ye[D * dy] += v[dy]; }
         }
      }
      if (VEC && c + 1 < ncomp) { __syncthreads(); } // F consumed before the next component's A
   }
}

// ---- persistent CTAs: every stream of the next tile is in flight while this tile computes ------
// One CTA per resident slot loops over tiles (tile += gridDim.x).  The one-shot kernel above pays,
// per tile, the lane -> (element, dx, dz) set-up and the dependent load chain gather index -> x ->
// stage 1 (65 % of the stall samples of the BP1 kernels, profiles/r01_bp1_p4_tuned_ncu.txt; with
// <= 3 resident CTAs at p >= 7 nothing hides it).  Here the set-up runs once per CTA and the chain
// is broken across tiles: the next tile's gather indices are loaded during stages 2-3, its x values
// during stages 4-5, and its quadrature-data slab is requested (TMA) the moment stage 3 has
// consumed the current one.  Scalar spaces only; same stages, same results as the one-shot kernel.
template <int KIND, int D, int Q, int NEB, bool FUSED, bool COLLOC, int MINB>
__global__ void __launch_bounds__(PipeCfg<KIND, D, Q, NEB>::NT, MINB)
pa_apply_3d_eo_persistent_kernel(const __grid_constant__ EOBasis<D, Q> bs, const double *__restrict__ qdata,
                                 const double *__restrict__ x, double *__restrict__ y,
                                 const int *__restrict__ idx, const int64_t ne, const int64_t e_first,
                                 const __grid_constant__ TmapArg<PipeCfg<KIND, D, Q, NEB>::SWZ> tmap)
{
   using C = PipeCfg<KIND, D, Q, NEB>;
   using S = EOStages<KIND, D, Q, NEB, COLLOC, false>;
   static_assert(!C::ALIAS, "the persistent form keeps the slab and the y-stage buffer apart");
   constexpr int D2 = C::D2, D3 = C::D3, Q3 = C::Q3, YQ = C::YQ;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   unsigned char *sbase = C::SWZ ? smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) : smem_raw;
   double *qd = reinterpret_cast<double *>(C::SWZ ? sbase : sbase + 16);
   double *r1 = qd + C::QD;
   double *r2 = r1 + C::R1;
   uint64_t *bar = reinterpret_cast<uint64_t *>(C::SWZ ? reinterpret_cast<unsigned char *>(r2 + C::R2) : sbase);

   const int t = threadIdx.x;
   const int64_t ntiles = (ne - e_first + NEB - 1) / NEB;
   int e1, dx1, dz1;
   eo_line_coords<D, NEB, C::L.order>(t, e1, dx1, dz1);
   const bool has1 = t < C::T1;
   const int loc1 = dx1 + D2 * dz1;
   const int a1 = Q * e1 + YQ * dx1 + C::SA * dz1;

   auto tile_nev = [&](int64_t tile) -> int
   {
      const int64_t left = ne - (e_first + tile * NEB);
      return (int)(left < NEB ? left : NEB);
   };
   auto issue_qdata = [&](int64_t tile)
   {
      const int64_t e0 = e_first + tile * NEB;
      const int nev = tile_nev(tile);
      const double *qslab = qdata + e0 * (int64_t)(C::NQD * Q3);
      if constexpr (C::TMA_OK)
      {
         if (t == 0)
         {
            fence_proxy_async(); // the slab was just read through the generic proxy
            mbar_expect_tx(bar, (uint32_t)(nev * C::NQD * Q3) * 8u);
            for (int e = 0; e < nev; e++)
            {
               if constexpr (C::SWZ)
               {
#pragma unroll
                  for (int b = 0; b < C::NBOX; b++)
                  {
                     tensor_g2s_3d(qd + e * C::DS + b * (C::BOX_ROWS * Q), &tmap, 0, b * C::BOX_ROWS, (int)(e0 + e), bar);
                  }
               }
               else
               {
                  bulk_g2s(qd + e * C::DS, qslab + (int64_t)e * (C::NQD * Q3), (uint32_t)(C::NQD * Q3) * 8u, bar);
               }
            }
         }
      }
      else
      {
         for (int i = t; i < nev * C::NQD * Q3; i += C::NT) { cp_async8(qd + i, qslab + i); }
      }
   };
   // gather indices of a tile (FUSED), -1 where this thread has no line or the dof is essential
   auto load_idx = [&](int64_t tile, int (&g)[D])
   {
#pragma unroll
      for (int dy = 0; dy < D; dy++) { g[dy] = -1; }
      if constexpr (FUSED)
      {
         if (has1 && e1 < tile_nev(tile))
         {
            const int *ip = idx + (e_first + tile * NEB + e1) * D3 + loc1;
#pragma unroll
            for (int dy = 0; dy < D; dy++) { g[dy] = __ldg(ip + D * dy); }
         }
      }
   };
   auto load_x = [&](int64_t tile, const int (&g)[D], double (&u)[D])
   {
#pragma unroll
      for (int dy = 0; dy < D; dy++) { u[dy] = 0.0; }
      if (has1 && e1 < tile_nev(tile))
      {
         if constexpr (FUSED)
         {
#pragma unroll
            for (int dy = 0; dy < D; dy++) { u[dy] = g[dy] >= 0 ? __ldg(x + g[dy]) : 0.0; }
         }
         else
         {
            const double *xe = x + (e_first + tile * NEB + e1) * D3 + loc1;
#pragma unroll
            for (int dy = 0; dy < D; dy++) { u[dy] = __ldg(xe + D * dy); }
         }
      }
   };

   int64_t tile = blockIdx.x; // the launcher keeps gridDim.x <= ntiles
   if (t == 0) { mbar_init(bar, 1); }
   __syncthreads();
   issue_qdata(tile);
   int gi[D], gi_n[D];
   double u[D], u_n[D];
   load_idx(tile, gi);
   load_x(tile, gi, u);
   uint32_t parity = 0;

   for (;;)
   {
      const int64_t next = tile + gridDim.x;
      const bool more = next < ntiles;
      const int nev = tile_nev(tile);
      if (has1) { S::stage1(bs, u, r1 + a1); }
      if (more) { load_idx(next, gi_n); }
      __syncthreads();
      if (t < C::T2) { S::stage2(bs, r1, r2, t); }
      if constexpr (!C::TMA_OK) { cp_async_wait_all(); }
      __syncthreads();
      if (t < C::T3) { S::stage3(bs, r2, qd, t, bar, parity); }
      parity ^= 1u;
      __syncthreads();
      // the slab is free: request the next tile's quadrature data and gather its inputs
      if (more)
      {
         issue_qdata(next);
         load_x(next, gi_n, u_n);
      }
      if (t < C::T2) { S::stage4(bs, r2, r1, t); }
      __syncthreads();
      if (has1 && e1 < nev)
      {
         double v[D];
         S::stage5(bs, r1 + a1, v);
         double *ye = y + (e_first + tile * NEB + e1) * D3 + loc1; // E-vector form
#pragma unroll
         for (int dy = 0; dy < D; dy++)
         {
            if (FUSED)
            {
               if (gi[dy] >= 0) { atomicAdd(y + gi[dy], v[dy]); }
            }
            else { ye[D * dy] += v[dy]; }
         }
      }
      if (!more) { break; }
#pragma unroll
      for (int dy = 0; dy < D; dy++) { gi[dy] = gi_n[dy]; u[dy] = u_n[dy]; }
      tile = next;
      __syncthreads(); // F consumed before the next tile's stage 1 overwrites the region
   }
}

// host: even-odd tables from the (Q x D, q fastest) tables; returns false when the basis does
// not have the required symmetry (relative 1e-12)
template <int D, int Q>
inline bool fill_eo_basis(const b2f_pa *pa, EOBasis<D, Q> &bs)
{
   constexpr int DH = (D + 1) / 2, QH = (Q + 1) / 2;
   double bmax = 0.0, gmax = 0.0, asym = 0.0;
   auto B = [&](int q, int d) { return pa->B[q + Q * d]; };
   auto G = [&](int q, int d) { return pa->G[q + Q * d]; };
   for (int q = 0; q < Q; q++)
      for (int d = 0; d < D; d++)
      {
         bmax = fmax(bmax, fabs(B(q, d)));
         gmax = fmax(gmax, fabs(G(q, d)));
      }
   for (int q = 0; q < Q; q++)
      for (int d = 0; d < D; d++)
      {
         asym = fmax(asym, fabs(B(q, d) - B(Q - 1 - q, D - 1 - d)) / (bmax > 0 ? bmax : 1));
         if (gmax > 0) { asym = fmax(asym, fabs(G(q, d) + G(Q - 1 - q, D - 1 - d)) / gmax); }
      }
   if (asym > 1e-12) { return false; }
   for (int q = 0; q < QH; q++)
      for (int i = 0; i < DH; i++)
      {
         const int j = D - 1 - i;
         // symmetrised entries
         const double bi = 0.5 * (B(q, i) + B(Q - 1 - q, j)), bj = 0.5 * (B(q, j) + B(Q - 1 - q, i));
         const double gi = 0.5 * (G(q, i) - G(Q - 1 - q, j)), gj = 0.5 * (G(q, j) - G(Q - 1 - q, i));
         if (i == j)
         {
            bs.Bs[q][i] = bi; bs.Ba[q][i] = 0.0;
            bs.Gs[q][i] = gi; bs.Ga[q][i] = 0.0;
         }
         else
         {
            bs.Bs[q][i] = 0.5 * (bi + bj); bs.Ba[q][i] = 0.5 * (bi - bj);
            bs.Gs[q][i] = 0.5 * (gi + gj); bs.Ga[q][i] = 0.5 * (gi - gj);
         }
      }
   return true;
}

// host: the (Q, rows-per-element, NE) tensor map over an integrator's quadrature data, boxes of
// (Q, box_rows, 1), 64-byte swizzle (Q = 8).  Cached in the b2f_pa (rebuilt when the data
// pointer or the element count changes: compacted copies of marked integrators, |D| twins).
inline int make_qdata_tensor_map(b2f_pa *pa, int q, int rows, int box_rows)  // q = 8: 64-byte swizzle, q = 4: 32-byte
{
   if (pa->tmap_data == pa->data && pa->tmap_ne == pa->ne) { return B2F_OK; }
   typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
   static encode_fn encode = nullptr;
   static std::mutex mtx;
   std::lock_guard<std::mutex> lock(mtx);
   if (!encode)
   {
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      B2F_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (!fn || qres != cudaDriverEntryPointSuccess)
      {
         set_error("cuTensorMapEncodeTiled is not available from this driver");
         return B2F_ERR_UNSUPPORTED;
      }
      encode = reinterpret_cast<encode_fn>(fn);
   }
   static_assert(sizeof(CUtensorMap) == sizeof(pa->tmap), "tensor map size");
   const cuuint64_t dims[3] = {(cuuint64_t)q, (cuuint64_t)rows, (cuuint64_t)pa->ne};
   const cuuint64_t strides[2] = {(cuuint64_t)q * 8u, (cuuint64_t)q * 8u * (cuuint64_t)rows};
   const cuuint32_t box[3] = {(cuuint32_t)q, (cuuint32_t)box_rows, 1u};
   const cuuint32_t estr[3] = {1u, 1u, 1u};
   const CUresult r = encode(reinterpret_cast<CUtensorMap *>(pa->tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                             pa->data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             q == 8 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
   if (r != CUDA_SUCCESS)
   {
      set_error("cuTensorMapEncodeTiled failed (%d) for q=%d rows=%d ne=%lld", (int)r, q, rows, (long long)pa->ne);
      return B2F_ERR_CUDA;
   }
   pa->tmap_data = pa->data;
   pa->tmap_ne = pa->ne;
   return B2F_OK;
}

// PERSIST: -1 = one-shot kernel only (persistent form not instantiated), 0 = one-shot by default,
// 1 = persistent by default (both instantiated; B2F_PERSIST overrides)
template <int KIND, int D, int Q, int NEB, int MINB = 1, int FUSEDSEL = 1, int PERSIST = -1>
int launch_3d_eo(b2f_pa *pa, const double *x, double *y, const int *idx, double *dot_acc = nullptr)
{
   using C = PipeCfg<KIND, D, Q, NEB>;
   b2f_ctx *ctx = pa->ctx;
   if ((int)C::SMEM > ctx->max_smem_optin)
   {
      set_error("pa_apply_3d_eo<D=%d,Q=%d,NEB=%d> needs %zu B of shared memory", D, Q, NEB, C::SMEM);
      return B2F_ERR_UNSUPPORTED;
   }
   EOBasis<D, Q> bs;
   if (!fill_eo_basis(pa, bs))
   {
      set_error("pa_apply_3d_eo: the basis tables are not symmetric");
      return B2F_ERR_UNSUPPORTED;
   }
   // element range of this launch (b2f_pa::rb / re; the whole mesh unless the operator splits
   // interior and boundary elements around the halo exchange)
   const int64_t e_first = pa->rb, e_last = pa->re < 0 ? pa->ne : pa->re;
   if (e_last <= e_first) { return B2F_OK; }
   const int64_t grid = cdiv(e_last - e_first, NEB);
   TmapArg<C::SWZ> tm;
   if constexpr (C::SWZ)
   {
      if (int rc = make_qdata_tensor_map(pa, Q, C::ROWS, C::BOX_ROWS)) { return rc; }
      memcpy(tm.bytes, pa->tmap, sizeof(tm.bytes));
   }
   auto go = [&](auto kern) -> int
   {
      int resident = 0;
      if (int rc = kernel_residency(ctx, kern, C::NT, C::SMEM, resident)) { return rc; }
      const char *pfs = getenv("B2F_PF");
      const int pf = pfs ? atoi(pfs) : resident;
      const VecArgs va = {pa->vd, (long long)pa->vxs, (long long)pa->vxo, (long long)pa->vics};
      timing_mark(ctx);
      kern<<<(unsigned)grid, C::NT, C::SMEM, ctx->stream>>>(bs, pa->data, x, y, idx, e_last, pf, dot_acc,
                                                           e_first, va, tm);
      timing_mark(ctx);
      B2F_LAUNCH_CHECK(ctx);
      return B2F_OK;
   };
   const bool fused = idx != nullptr;
   const bool vec = pa->vd > 1;
   // persistent form (scalar spaces, no fused dot): tuned per (kind, D, Q), B2F_PERSIST=0|1 overrides
   if constexpr (PERSIST >= 0)
   {
      const char *ps = getenv("B2F_PERSIST");
      const bool persist = ps ? atoi(ps) != 0 : PERSIST > 0;
      if (persist && !vec && !dot_acc && ((fused && FUSEDSEL != 0) || (!fused && FUSEDSEL != 1)))
      {
         auto gop = [&](auto kern) -> int
         {
            int resident = 0;
            if (int rc = kernel_residency(ctx, kern, C::NT, C::SMEM, resident)) { return rc; }
            const int64_t g = grid < resident ? grid : resident;
            timing_mark(ctx);
            kern<<<(unsigned)g, C::NT, C::SMEM, ctx->stream>>>(bs, pa->data, x, y, idx, e_last, e_first, tm);
            timing_mark(ctx);
            B2F_LAUNCH_CHECK(ctx);
            return B2F_OK;
         };
         constexpr bool CLP = (D == Q);
         if (pa->collocated && CLP)
         {
            if constexpr (FUSEDSEL != 0) { if (fused) { return gop(pa_apply_3d_eo_persistent_kernel<KIND, D, Q, NEB, true, CLP, MINB>); } }
            if constexpr (FUSEDSEL != 1) { if (!fused) { return gop(pa_apply_3d_eo_persistent_kernel<KIND, D, Q, NEB, false, CLP, MINB>); } }
         }
         if constexpr (FUSEDSEL != 0) { if (fused) { return gop(pa_apply_3d_eo_persistent_kernel<KIND, D, Q, NEB, true, false, MINB>); } }
         if constexpr (FUSEDSEL != 1) { if (!fused) { return gop(pa_apply_3d_eo_persistent_kernel<KIND, D, Q, NEB, false, false, MINB>); } }
      }
   }
   if ((fused && FUSEDSEL == 0) || (!fused && FUSEDSEL == 1) || (dot_acc && (!fused || vec)))
   {
      set_error("pa_apply_3d_eo: kernel form not instantiated");
      return B2F_ERR_UNSUPPORTED;
   }
   constexpr bool CL = (D == Q);
   // (the vector form has its own instantiation so that the scalar hot path is untouched)
   if (pa->collocated && CL)
   {
      if constexpr (FUSEDSEL != 0)
      {
         if (fused && vec) { return go(pa_apply_3d_eo_kernel<KIND, D, Q, NEB, true, CL, MINB, false, true>); }
         if (fused && dot_acc) { return go(pa_apply_3d_eo_kernel<KIND, D, Q, NEB, true, CL, MINB, true>); }
         if (fused) { return go(pa_apply_3d_eo_kernel<KIND, D, Q, NEB, true, CL, MINB>); }
      }
      if constexpr (FUSEDSEL != 1)
      {
         if (!fused && vec) { return go(pa_apply_3d_eo_kernel<KIND, D, Q, NEB, false, CL, MINB, false, true>); }
         if (!fused) { return go(pa_apply_3d_eo_kernel<KIND, D, Q, NEB, false, CL, MINB>); }
      }
   }
   if constexpr (FUSEDSEL != 0)
   {
      if (fused && vec) { return go(pa_apply_3d_eo_kernel<KIND, D, Q, NEB, true, false, MINB, false, true>); }
      if (fused && dot_acc) { return go(pa_apply_3d_eo_kernel<KIND, D, Q, NEB, true, false, MINB, true>); }
      if (fused) { return go(pa_apply_3d_eo_kernel<KIND, D, Q, NEB, true, false, MINB>); }
   }
   if constexpr (FUSEDSEL != 1)
   {
      if (!fused && vec) { return go(pa_apply_3d_eo_kernel<KIND, D, Q, NEB, false, false, MINB, false, true>); }
      if (!fused) { return go(pa_apply_3d_eo_kernel<KIND, D, Q, NEB, false, false, MINB>); }
   }
   return B2F_ERR_UNSUPPORTED;
}

} // namespace b2f
