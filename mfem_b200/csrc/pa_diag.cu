// mfem_b200/csrc/pa_diag.cu -- AssembleDiagonalPA (setup path, run once per solve).
//
// Computes what (Smem)PADiffusionDiagonal2D/3D (fem/integ/bilininteg_diffusion_kernels.hpp:
// 92-485) and (Smem)PAMassAssembleDiagonal2D/3D (fem/integ/bilininteg_mass_kernels.hpp:132-416)
// compute -- diag_E(d) = sum_q sum_ij D_ij(q) dphi_i(d,q) dphi_j(d,q) -- with a formulation of
// its own, built for one pass over the quadrature data:
//
//  * The tensor basis makes every product dphi_i dphi_j a product of three 1D factors taken
//    from the SQUARED tables  BB = B.B,  GG = G.G,  BG = B.G  (Q x D each):
//        11 -> GGx BBy BBz   22 -> BBx GGy BBz   33 -> BBx BBy GGz
//        12 -> BGx BGy BBz   13 -> BGx BBy BGz   23 -> BBx BGy BGz      (off-diagonals count twice)
//    so the diagonal is three sum-factorised contractions (x, then y, then z) of the six
//    quadrature-data components at once.  After the y-contraction the components that share
//    their z-table are summed ({11,12,22} -> BBz, {13,23} -> BGz, {33} -> GGz): 6 fields after x,
//    3 after y, 1 after z.  The mass diagonal is the single chain BB BB BB.
//  * Pencil decomposition as in the apply kernel (pa_apply_eo.cuh): one thread owns one 1D line
//    per stage.  In the x-stage a thread reads its six rows of D straight from global memory
//    with 128-bit loads (consecutive threads read consecutive rows: the CTA streams its elements'
//    contiguous slabs exactly once), so the quadrature data crosses HBM once and nothing else
//    does: the kernel is bound by the 24 GB of D at 151 M dofs (3.8 ms at the measured peak).
//  * FUSED: the result is scattered straight into the L-vector with unsigned FP64 reductions
//    (ElementRestriction::AbsMultTranspose, fem/restriction.cpp:203-228), the E-vector diagonal
//    never exists.  The E-vector form (the AssembleDiagonalPA seam) accumulates into diag_E.
//
// Specialised for every (D1D, Q1D) of the apply kernels' dispatch table; other sizes (and 2D)
// run the run-time-sized kernels at the end of the file (same formulation, tables in shared
// memory).
#include "common.cuh"

namespace b2f
{

template <int D, int Q>
struct DiagTables
{
   double BB[Q][D], GG[Q][D], BG[Q][D];
};

template <int KIND, int D, int Q, int NEB>
struct DiagCfg
{
   static constexpr bool DIFF = (KIND == B2F_KIND_DIFFUSION);
   static constexpr int NX = DIFF ? 6 : 1; // fields after the x-contraction
   static constexpr int NY = DIFF ? 3 : 1; // fields after the y-contraction
   static constexpr int YQ = Q * NEB;      // yq = qy + Q e, zq = qz + Q e
   // a 1D line of Q values is stored with an ODD pitch: the next stage's lanes each read one line,
   // the same position of every line per instruction, i.e. addresses QP apart -> distinct banks
   static constexpr int QP = Q | 1;
   static constexpr int T1 = NEB * Q * Q, T2 = NEB * D * Q, T3 = NEB * D * D;
   static constexpr int NT = ((T1 + 31) / 32) * 32;
   static constexpr int XS = NX * D * Q * NEB * QP; // [k][dx][qz][e][qy (pitch QP)]: lines over qy
   static constexpr int YS = NY * NEB * D * D * QP; // [f][e][dy][dx][qz (pitch QP)]: lines over qz
   static constexpr size_t SMEM = sizeof(double) * (size_t)(XS + YS);
};

constexpr int diag_neb(int kind, int D, int Q)
{
   // elements per CTA: <= 256 threads, <= 40 KB of shared memory
   int n = 1;
   while (n < 16)
   {
      const int m = n + 1, qp = Q | 1;
      const int nx = kind == B2F_KIND_DIFFUSION ? 6 : 1, ny = kind == B2F_KIND_DIFFUSION ? 3 : 1;
      if (m * Q * Q > 256 || 8 * (nx * D * Q * m * qp + ny * m * D * D * qp) > 40 * 1024) { break; }
      n = m;
   }
   return n;
}

template <int Q>
__device__ __forceinline__ void ldg_row(const double *__restrict__ p, double (&v)[Q])
{
   if constexpr (Q % 2 == 0)
   {
      const double2 *p2 = reinterpret_cast<const double2 *>(p);
#pragma unroll
      for (int i = 0; i < Q / 2; i++) { const double2 w = __ldg(p2 + i); v[2 * i] = w.x; v[2 * i + 1] = w.y; }
   }
   else
   {
#pragma unroll
      for (int i = 0; i < Q; i++) { v[i] = __ldg(p + i); }
   }
}

// NCOMP: stored components per point: 1 (mass), 6 (symmetric) or 9 (full matrix, row-major)
template <int KIND, int D, int Q, int NCOMP, int NEB, bool FUSED>
__global__ void __launch_bounds__(DiagCfg<KIND, D, Q, NEB>::NT)
pa_diag_3d_kernel(const __grid_constant__ DiagTables<D, Q> tb, const double *__restrict__ qdata,
                  double *__restrict__ out, const int *__restrict__ gmap, const int64_t ne)
{
   using C = DiagCfg<KIND, D, Q, NEB>;
   constexpr int Q3 = Q * Q * Q, D3 = D * D * D, QP = C::QP, LX = NEB * QP;
   extern __shared__ __align__(16) double dsm[];
   double *Xs = dsm, *Ys = dsm + C::XS;
   const int t = threadIdx.x;
   const int64_t e0 = (int64_t)blockIdx.x * NEB;
   const int nev = (int)((ne - e0) < NEB ? (ne - e0) : NEB);

   // ---- x: line (qy, qz, e), the NCOMP rows of D over qx -> 6 (1) fields at dx
   if (t < C::T1)
   {
      const int yq = t % C::YQ, qz = t / C::YQ;
      const int e = yq / Q, qy = yq - e * Q;
      double r[C::NX][Q];
      if (e < nev)
      {
         const double *row = qdata + (e0 + e) * (int64_t)(NCOMP * Q3) + Q * (qy + Q * qz);
         if constexpr (!C::DIFF) { ldg_row<Q>(row, r[0]); }
         else if constexpr (NCOMP == 6)
         {
            // (11, 21, 31, 22, 32, 33): off-diagonals enter twice
#pragma unroll
            for (int k = 0; k < 6; k++) { ldg_row<Q>(row + k * Q3, r[k]); }
#pragma unroll
            for (int q = 0; q < Q; q++) { r[1][q] += r[1][q]; r[2][q] += r[2][q]; r[4][q] += r[4][q]; }
         }
         else
         {
            // full matrix, row-major i*3+j: (11, 12+21, 13+31, 22, 23+32, 33)
            double s[Q];
            ldg_row<Q>(row + 0 * Q3, r[0]);
            ldg_row<Q>(row + 1 * Q3, r[1]); ldg_row<Q>(row + 3 * Q3, s);
#pragma unroll
            for (int q = 0; q < Q; q++) { r[1][q] += s[q]; }
            ldg_row<Q>(row + 2 * Q3, r[2]); ldg_row<Q>(row + 6 * Q3, s);
#pragma unroll
            for (int q = 0; q < Q; q++) { r[2][q] += s[q]; }
            ldg_row<Q>(row + 4 * Q3, r[3]);
            ldg_row<Q>(row + 5 * Q3, r[4]); ldg_row<Q>(row + 7 * Q3, s);
#pragma unroll
            for (int q = 0; q < Q; q++) { r[4][q] += s[q]; }
            ldg_row<Q>(row + 8 * Q3, r[5]);
         }
      }
      else
      {
#pragma unroll
         for (int k = 0; k < C::NX; k++)
#pragma unroll
            for (int q = 0; q < Q; q++) { r[k][q] = 0.0; }
      }
      double *X0 = Xs + qz * LX + e * QP + qy;
#pragma unroll
      for (int dx = 0; dx < D; dx++)
      {
         double a[C::NX];
#pragma unroll
         for (int k = 0; k < C::NX; k++) { a[k] = 0.0; }
#pragma unroll
         for (int qx = 0; qx < Q; qx++)
         {
            const double bb = tb.BB[qx][dx];
            if constexpr (C::DIFF)
            {
               const double gg = tb.GG[qx][dx], bg = tb.BG[qx][dx];
               a[0] = fma(gg, r[0][qx], a[0]); // 11
               a[1] = fma(bg, r[1][qx], a[1]); // 12
               a[2] = fma(bg, r[2][qx], a[2]); // 13
               a[3] = fma(bb, r[3][qx], a[3]); // 22
               a[4] = fma(bb, r[4][qx], a[4]); // 23
               a[5] = fma(bb, r[5][qx], a[5]); // 33
            }
            else { a[0] = fma(bb, r[0][qx], a[0]); }
         }
#pragma unroll
         for (int k = 0; k < C::NX; k++) { X0[(k * D + dx) * (Q * LX)] = a[k]; }
      }
   }
   __syncthreads();
   // ---- y: line (qz, e, dx) over qy -> 3 (1) fields at dy, grouped by their z-table
   if (t < C::T2)
   {
      const int zq = t % C::YQ, dx = t / C::YQ; // zq = qz + Q e
      const int e = zq / Q, qz = zq - e * Q;
      double r[C::NX][Q];
#pragma unroll
      for (int k = 0; k < C::NX; k++)
#pragma unroll
         for (int qy = 0; qy < Q; qy++) { r[k][qy] = Xs[((k * D + dx) * Q + qz) * LX + e * QP + qy]; }
      double *Y0 = Ys + (e * D * D + dx) * QP + qz; // + dy * D * QP, + f * NEB * D * D * QP
#pragma unroll
      for (int dy = 0; dy < D; dy++)
      {
         double zb = 0.0, zbg = 0.0, zg = 0.0;
#pragma unroll
         for (int qy = 0; qy < Q; qy++)
         {
            const double bb = tb.BB[qy][dy];
            if constexpr (C::DIFF)
            {
               const double gg = tb.GG[qy][dy], bg = tb.BG[qy][dy];
               zb = fma(bb, r[0][qy], fma(bg, r[1][qy], fma(gg, r[3][qy], zb))); // 11, 12, 22 -> BBz
               zbg = fma(bb, r[2][qy], fma(bg, r[4][qy], zbg));                  // 13, 23     -> BGz
               zg = fma(bb, r[5][qy], zg);                                       // 33         -> GGz
            }
            else { zb = fma(bb, r[0][qy], zb); }
         }
         Y0[dy * (D * QP)] = zb;
         if constexpr (C::DIFF)
         {
            Y0[NEB * D * D * QP + dy * (D * QP)] = zbg;
            Y0[2 * NEB * D * D * QP + dy * (D * QP)] = zg;
         }
      }
   }
   __syncthreads();
   // ---- z: line (dx, dy, e) over qz -> the diagonal at dz
   if (t < C::T3)
   {
      const int dx = t % D, dy = (t / D) % D, e = t / (D * D);
      if (e < nev)
      {
         const int64_t base = (e0 + e) * (int64_t)D3 + dx + D * dy;
         int g[D]; // scatter targets first: their latency hides behind the shared-memory reads
         if constexpr (FUSED)
         {
#pragma unroll
            for (int dz = 0; dz < D; dz++) { g[dz] = __ldg(gmap + base + (int64_t)(D * D) * dz); }
         }
         double r[C::NY][Q];
#pragma unroll
         for (int f = 0; f < C::NY; f++)
#pragma unroll
            for (int qz = 0; qz < Q; qz++) { r[f][qz] = Ys[((f * NEB + e) * D * D + dy * D + dx) * QP + qz]; }
#pragma unroll
         for (int dz = 0; dz < D; dz++)
         {
            double s = 0.0;
#pragma unroll
            for (int qz = 0; qz < Q; qz++)
            {
               s = fma(tb.BB[qz][dz], r[0][qz], s);
               if constexpr (C::DIFF) { s = fma(tb.BG[qz][dz], r[1][qz], fma(tb.GG[qz][dz], r[2][qz], s)); }
            }
            const int64_t i = base + (int64_t)(D * D) * dz;
            if constexpr (FUSED) { atomicAdd(out + (g[dz] >= 0 ? g[dz] : -1 - g[dz]), s); } // unsigned: AbsMultTranspose
            else { out[i] += s; }
         }
      }
   }
}

template <int D, int Q>
static void fill_diag_tables(const b2f_pa *pa, DiagTables<D, Q> &tb)
{
   for (int q = 0; q < Q; q++)
      for (int d = 0; d < D; d++)
      {
         const double b = pa->B[q + Q * d], g = pa->G[q + Q * d];
         tb.BB[q][d] = b * b; tb.GG[q][d] = g * g; tb.BG[q][d] = b * g;
      }
}

template <int KIND, int D, int Q>
static int launch_diag_3d(b2f_pa *pa, double *out, const int *gmap)
{
   constexpr int NEB = diag_neb(KIND, D, Q);
   using C = DiagCfg<KIND, D, Q, NEB>;
   b2f_ctx *ctx = pa->ctx;
   DiagTables<D, Q> tb;
   fill_diag_tables(pa, tb);
   const unsigned grid = (unsigned)cdiv(pa->ne, NEB);
   auto go = [&](auto kern) -> int
   {
      int resident = 0;
      if (int rc = kernel_residency(ctx, kern, C::NT, C::SMEM, resident)) { return rc; }
      kern<<<grid, C::NT, C::SMEM, ctx->stream>>>(tb, pa->data, out, gmap, pa->ne);
      B2F_LAUNCH_CHECK(ctx);
      return B2F_OK;
   };
   if constexpr (KIND == B2F_KIND_MASS)
   {
      return gmap ? go(pa_diag_3d_kernel<KIND, D, Q, 1, NEB, true>) : go(pa_diag_3d_kernel<KIND, D, Q, 1, NEB, false>);
   }
   else
   {
      if (pa->ncomp == 9)
      {
         return gmap ? go(pa_diag_3d_kernel<KIND, D, Q, 9, NEB, true>) : go(pa_diag_3d_kernel<KIND, D, Q, 9, NEB, false>);
      }
      return gmap ? go(pa_diag_3d_kernel<KIND, D, Q, 6, NEB, true>) : go(pa_diag_3d_kernel<KIND, D, Q, 6, NEB, false>);
   }
}

// ---- run-time sized kernels: 2D, and 3D sizes outside the dispatch table -----------------------
// Same formulation, one CTA per element, squared tables and the intermediate fields in shared
// memory.  To keep the footprint at 3 x-fields + 1 y-field (D1D, Q1D up to 16 fit the opt-in
// limit) the 3D kernel makes one pass per z-table group: {11, 12, 22} -> BBz, {13, 23} -> BGz,
// {33} -> GGz; every output dof is owned by one thread across the passes.
__global__ void __launch_bounds__(128)
pa_diag_generic_kernel(int kind, int dim, int D, int Q, int ncomp, const double *__restrict__ Bg,
                       const double *__restrict__ Gg, const double *__restrict__ qdata,
                       double *__restrict__ out, const int *__restrict__ gmap)
{
   extern __shared__ __align__(16) double dsm[];
   const bool diff = kind == B2F_KIND_DIFFUSION;
   const int QD = Q * D;
   double *BB = dsm, *GG = BB + QD, *BG = GG + QD, *Xs = BG + QD;
   const int nq = dim == 3 ? Q * Q * Q : Q * Q, nd = dim == 3 ? D * D * D : D * D;
   const int lines = dim == 3 ? Q * Q : Q; // x-lines per field
   const int xf = lines * D;               // one x-field: [line][dx]
   double *Ys = Xs + 3 * xf;               // one y-field (3D): [qz][dy][dx]
   const int64_t e = blockIdx.x;
   const double *De = qdata + e * (int64_t)ncomp * nq;
   for (int i = threadIdx.x; i < QD; i += blockDim.x)
   {
      const double b = Bg[i], g = diff ? Gg[i] : 0.0;
      BB[i] = b * b; GG[i] = g * g; BG[i] = b * g;
   }
   // symmetric-sum component s of the stored data at point q: 3D (11, 12+21, 13+31, 22, 23+32, 33),
   // 2D (11, 12+21, 22); storage orders as in the apply kernels (symmetric: 11,21,31,22,32,33 /
   // 11,21,22; full: row-major)
   auto comp = [&](int s, int q) -> double
   {
      if (!diff) { return De[q]; }
      if (dim == 3)
      {
         if (ncomp == 9)
         {
            switch (s)
            {
               case 0: return De[q];
               case 1: return De[q + 1 * nq] + De[q + 3 * nq];
               case 2: return De[q + 2 * nq] + De[q + 6 * nq];
               case 3: return De[q + 4 * nq];
               case 4: return De[q + 5 * nq] + De[q + 7 * nq];
               default: return De[q + 8 * nq];
            }
         }
         const double v = De[q + s * nq];
         return (s == 1 || s == 2 || s == 4) ? v + v : v;
      }
      if (ncomp == 4) { return s == 0 ? De[q] : s == 1 ? De[q + nq] + De[q + 2 * nq] : De[q + 3 * nq]; }
      const double v = De[q + s * nq];
      return s == 1 ? v + v : v;
   };
   auto emit = [&](int w, double a)
   {
      const int64_t i = e * nd + w;
      if (gmap) { const int g = gmap[i]; atomicAdd(out + (g >= 0 ? g : -1 - g), a); }
      else { out[i] += a; }
   };
   // x-contraction of component s with table T into x-field slot
   auto xstage = [&](int slot, int s, const double *T)
   {
      for (int w = threadIdx.x; w < xf; w += blockDim.x)
      {
         const int dx = w % D, l = w / D;
         double a = 0.0;
         for (int qx = 0; qx < Q; qx++) { a = fma(T[qx + Q * dx], comp(s, qx + Q * l), a); }
         Xs[slot * xf + w] = a;
      }
   };
   __syncthreads();
   if (dim == 2)
   {
      // 11 -> GGx BBy, 12 -> BGx BGy, 22 -> BBx GGy; mass BBx BBy
      xstage(0, 0, diff ? GG : BB);
      if (diff) { xstage(1, 1, BG); xstage(2, 2, BB); }
      __syncthreads();
      for (int w = threadIdx.x; w < nd; w += blockDim.x)
      {
         const int dx = w % D, dy = w / D;
         double a = 0.0;
         for (int qy = 0; qy < Q; qy++)
         {
            a = fma(BB[qy + Q * dy], Xs[qy * D + dx], a);
            if (diff)
            {
               a = fma(BG[qy + Q * dy], Xs[xf + qy * D + dx], a);
               a = fma(GG[qy + Q * dy], Xs[2 * xf + qy * D + dx], a);
            }
         }
         emit(w, a);
      }
      return;
   }
   const int ngroups = diff ? 3 : 1;
   for (int grp = 0; grp < ngroups; grp++)
   {
      // group 0: (11: GGx BBy), (12: BGx BGy), (22: BBx GGy) -> BBz
      // group 1: (13: BGx BBy), (23: BBx BGy)                -> BGz
      // group 2: (33: BBx BBy)                               -> GGz
      int nf = 1;
      const double *ytab[3] = {BB, BB, BB};
      if (!diff) { xstage(0, 0, BB); }
      else if (grp == 0) { xstage(0, 0, GG); xstage(1, 1, BG); xstage(2, 3, BB); ytab[1] = BG; ytab[2] = GG; nf = 3; }
      else if (grp == 1) { xstage(0, 2, BG); xstage(1, 4, BB); ytab[1] = BG; nf = 2; }
      else { xstage(0, 5, BB); }
      const double *ztab = grp == 0 ? BB : grp == 1 ? BG : GG;
      __syncthreads();
      for (int w = threadIdx.x; w < Q * D * D; w += blockDim.x)
      {
         const int dx = w % D, dy = (w / D) % D, qz = w / (D * D);
         double a = 0.0;
         for (int f = 0; f < nf; f++)
         {
            const double *T = ytab[f], *X = Xs + f * xf;
            for (int qy = 0; qy < Q; qy++) { a = fma(T[qy + Q * dy], X[(qz * Q + qy) * D + dx], a); }
         }
         Ys[w] = a;
      }
      __syncthreads();
      for (int w = threadIdx.x; w < nd; w += blockDim.x)
      {
         const int dxy = w % (D * D), dz = w / (D * D);
         double a = 0.0;
         for (int qz = 0; qz < Q; qz++) { a = fma(ztab[qz + Q * dz], Ys[qz * D * D + dxy], a); }
         emit(w, a);
      }
      __syncthreads(); // Xs, Ys are rewritten by the next group
   }
}

static int launch_diag_generic(b2f_pa *pa, double *out, const int *gmap)
{
   b2f_ctx *ctx = pa->ctx;
   const int D = pa->d1d, Q = pa->q1d;
   const size_t smem = sizeof(double) * ((size_t)3 * Q * D + (size_t)3 * (pa->dim == 3 ? Q * Q * D : Q * D) +
                                         (pa->dim == 3 ? (size_t)Q * D * D : 0));
   if ((int)smem > ctx->max_smem_optin)
   {
      set_error("pa_diag: D1D=%d Q1D=%d needs %zu B of shared memory (device limit %d)", D, Q, smem,
                ctx->max_smem_optin);
      return B2F_ERR_UNSUPPORTED;
   }
   if (smem > 48 * 1024)
   {
      B2F_CUDA(cudaFuncSetAttribute(pa_diag_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   }
   pa_diag_generic_kernel<<<(unsigned)pa->ne, 128, smem, ctx->stream>>>(pa->kind, pa->dim, D, Q, pa->ncomp, pa->d_B,
                                                                       pa->d_G, pa->data, out, gmap);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

#define B2F_DIAG_CASE(D_, Q_)                                                                    \
   case (D_) * 100 + (Q_):                                                                       \
      return pa->kind == B2F_KIND_MASS ? launch_diag_3d<B2F_KIND_MASS, D_, Q_>(pa, out, gmap)     \
                                       : launch_diag_3d<B2F_KIND_DIFFUSION, D_, Q_>(pa, out, gmap);
#define B2F_DIAG_ROW(D_) B2F_DIAG_CASE(D_, D_) B2F_DIAG_CASE(D_, D_ + 1) B2F_DIAG_CASE(D_, D_ + 2)

// E-vector form (gmap == nullptr): diag_e += diagonal; fused form: out[|gmap[i]|] += diagonal
static int diag_dispatch(b2f_pa *pa, double *out, const int *gmap)
{
   if (pa->ne == 0) { return B2F_OK; }
   if (pa->dim == 3 && !getenv("B2F_GENERAL_KERNEL"))
   {
      switch (pa->d1d * 100 + pa->q1d)
      {
         B2F_DIAG_ROW(2) B2F_DIAG_ROW(3) B2F_DIAG_ROW(4) B2F_DIAG_ROW(5)
         B2F_DIAG_ROW(6) B2F_DIAG_ROW(7) B2F_DIAG_ROW(8) B2F_DIAG_ROW(9)
         default: break;
      }
   }
   return launch_diag_generic(pa, out, gmap);
}

int pa_diag(b2f_pa *pa, double *diag_e) { return diag_dispatch(pa, diag_e, nullptr); }

int pa_diag_fused(b2f_pa *pa, const int *gather_map, double *diag_l) { return diag_dispatch(pa, diag_l, gather_map); }

} // namespace b2f
