// mfem_b200/csrc/pa_apply_generic.cu -- run-time (D1D, Q1D) sum-factorised PA apply kernels.
//
// The role of MFEM's non-specialised PADiffusionApply2D/3D and PAMassApply2D/3D
// (fem/integ/bilininteg_diffusion_kernels.hpp:514-633,796-986, bilininteg_mass_kernels.hpp:
// 440-536,664-813): whatever the specialised kernels do not cover --
//   * NON-SYMMETRIC matrix coefficients (dim*dim quadrature-data components, row-major
//     11,12,13,21,... as PADiffusionSetup stores them),
//   * (D1D, Q1D) outside the dispatch tables (Q1D > D1D+2, D1D > 9, ...), up to what fits in
//     shared memory.
// One CTA per element, every intermediate in shared memory, each stage a strided loop over
// its output index space; E-vector or fused L-vector form (gather index, -1 = constrained dof,
// atomic scatter).  Correctness path, not a hot path: unlike the reference's fallback (one
// THREAD per element with ~76 KB of local memory) it still runs a full CTA per element.
#include "common.cuh"

namespace b2f
{

// diffusion D-apply at one point: v = O u with O symmetric (6 / 3 comps) or full (9 / 4)
template <int DIM>
__device__ __forceinline__ void apply_qdata(const double *__restrict__ Dq, int64_t stride, int ncomp,
                                            const double *u, double *v)
{
   if (DIM == 3)
   {
      if (ncomp == 6)
      {
         const double O11 = Dq[0], O12 = Dq[stride], O13 = Dq[2 * stride];
         const double O22 = Dq[3 * stride], O23 = Dq[4 * stride], O33 = Dq[5 * stride];
         v[0] = (O11 * u[0]) + (O12 * u[1]) + (O13 * u[2]);
         v[1] = (O12 * u[0]) + (O22 * u[1]) + (O23 * u[2]);
         v[2] = (O13 * u[0]) + (O23 * u[1]) + (O33 * u[2]);
      }
      else
      {
         for (int i = 0; i < 3; i++)
         {
            v[i] = (Dq[(3 * i) * stride] * u[0]) + (Dq[(3 * i + 1) * stride] * u[1]) +
                   (Dq[(3 * i + 2) * stride] * u[2]);
         }
      }
   }
   else
   {
      if (ncomp == 3)
      {
         const double O11 = Dq[0], O12 = Dq[stride], O22 = Dq[2 * stride];
         v[0] = (O11 * u[0]) + (O12 * u[1]);
         v[1] = (O12 * u[0]) + (O22 * u[1]);
      }
      else
      {
         // 2D stores (11, 21, 12, 22) (bilininteg_diffusion_kernels.hpp:590-593)
         v[0] = (Dq[0] * u[0]) + (Dq[2 * stride] * u[1]);
         v[1] = (Dq[stride] * u[0]) + (Dq[3 * stride] * u[1]);
      }
   }
}

// ---- 3D -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pa_apply_generic_3d_kernel(int kind, int D, int Q, int ncomp, const double *__restrict__ Bg,
                           const double *__restrict__ Gg, const double *__restrict__ qdata,
                           const double *__restrict__ x, double *__restrict__ y,
                           const int *__restrict__ idx, int vd, long long vxs, long long vxo, long long vics)
{
   // vector spaces (vdim > 1): blockIdx.y = component; component c of L-dof g is x[g*vxs + c*vxo],
   // its gather index starts at idx + c*vics, its E-vector block is (e*vd + c)
   const int comp = blockIdx.y;
   extern __shared__ double sm[];
   const bool diff = kind == B2F_KIND_DIFFUSION;
   const int D2 = D * D, D3 = D2 * D, Q2 = Q * Q, Q3 = Q2 * Q;
   const int nf1 = diff ? 2 : 1, nf2 = diff ? 3 : 1; // fields after the x / y (and z) stages
   double *sB = sm, *sG = sB + Q * D;
   double *X = sG + Q * D;              // [dz][dy][dx]
   double *A = X + D3;                  // [f][dz][dy][qx]      f: B_x u, G_x u
   double *C = A + nf1 * D2 * Q;        // [f][dz][qy][qx]      f: BB, GB (d/dx), BG (d/dy)
   double *E = C + nf2 * D * Q2;        // [f][qz][qy][qx]      gradient (or value) at the points
   const int64_t e = blockIdx.x;
   const int t = threadIdx.x, nt = blockDim.x;
   for (int i = t; i < Q * D; i += nt) { sB[i] = Bg[i]; sG[i] = Gg ? Gg[i] : 0.0; }
   const int *ix = idx ? idx + comp * vics + e * D3 : nullptr;
   const int64_t eoff = (e * vd + comp) * (int64_t)D3;
   for (int i = t; i < D3; i += nt)
   {
      if (idx) { const int g = ix[i]; X[i] = g >= 0 ? x[g * vxs + comp * vxo] : 0.0; }
      else { X[i] = x[eoff + i]; }
   }
   __syncthreads();
   // x
   for (int o = t; o < D2 * Q; o += nt)
   {
      const int qx = o % Q, r = o / Q; // r = dy + D*dz
      double vb = 0.0, vg = 0.0;
      for (int dx = 0; dx < D; dx++)
      {
         const double u = X[dx + D * r];
         vb = fma(sB[qx + Q * dx], u, vb);
         vg = fma(sG[qx + Q * dx], u, vg);
      }
      A[o] = vb;
      if (diff) { A[D2 * Q + o] = vg; }
   }
   __syncthreads();
   // y
   for (int o = t; o < D * Q2; o += nt)
   {
      const int qx = o % Q, qy = (o / Q) % Q, dz = o / Q2;
      double bb = 0.0, gb = 0.0, bg = 0.0;
      for (int dy = 0; dy < D; dy++)
      {
         const int a = qx + Q * (dy + D * dz);
         bb = fma(sB[qy + Q * dy], A[a], bb);
         if (diff)
         {
            gb = fma(sB[qy + Q * dy], A[D2 * Q + a], gb);
            bg = fma(sG[qy + Q * dy], A[a], bg);
         }
      }
      C[o] = bb;
      if (diff) { C[D * Q2 + o] = gb; C[2 * D * Q2 + o] = bg; }
   }
   __syncthreads();
   // z + quadrature data
   const double *De = qdata + e * (int64_t)ncomp * Q3;
   for (int o = t; o < Q3; o += nt)
   {
      const int r = o % Q2, qz = o / Q2; // r = qx + Q*qy
      double u[3] = {0.0, 0.0, 0.0}, v[3];
      for (int dz = 0; dz < D; dz++)
      {
         const int c = r + Q2 * dz;
         if (diff)
         {
            u[0] = fma(sB[qz + Q * dz], C[D * Q2 + c], u[0]);     // d/dx
            u[1] = fma(sB[qz + Q * dz], C[2 * D * Q2 + c], u[1]); // d/dy
            u[2] = fma(sG[qz + Q * dz], C[c], u[2]);              // d/dz
         }
         else { u[0] = fma(sB[qz + Q * dz], C[c], u[0]); }
      }
      if (diff)
      {
         apply_qdata<3>(De + o, Q3, ncomp, u, v);
         E[o] = v[0]; E[Q3 + o] = v[1]; E[2 * Q3 + o] = v[2];
      }
      else { E[o] = De[o] * u[0]; }
   }
   __syncthreads();
   // z^T : C[f][dz][qy][qx]   f: (for B_y B_x... ) reuse the same field meaning transposed
   for (int o = t; o < D * Q2; o += nt)
   {
      const int r = o % Q2, dz = o / Q2;
      double c0 = 0.0, c1 = 0.0, c2 = 0.0;
      for (int qz = 0; qz < Q; qz++)
      {
         const int q = r + Q2 * qz;
         if (diff)
         {
            c0 = fma(sB[qz + Q * dz], E[q], c0);          // from d/dx
            c1 = fma(sB[qz + Q * dz], E[Q3 + q], c1);     // from d/dy
            c2 = fma(sG[qz + Q * dz], E[2 * Q3 + q], c2); // from d/dz
         }
         else { c0 = fma(sB[qz + Q * dz], E[q], c0); }
      }
      C[o] = c0;
      if (diff) { C[D * Q2 + o] = c1; C[2 * D * Q2 + o] = c2; }
   }
   __syncthreads();
   // y^T : A[f][dz][dy][qx]   f0 pairs with G_x^T, f1 with B_x^T
   for (int o = t; o < D2 * Q; o += nt)
   {
      const int qx = o % Q, dy = (o / Q) % D, dz = o / (Q * D);
      double a0 = 0.0, a1 = 0.0;
      for (int qy = 0; qy < Q; qy++)
      {
         const int c = qx + Q * (qy + Q * dz);
         if (diff)
         {
            a0 = fma(sB[qy + Q * dy], C[c], a0);                 // -> G_x^T
            a1 = fma(sG[qy + Q * dy], C[D * Q2 + c], a1);        // d/dy part -> B_x^T
            a1 = fma(sB[qy + Q * dy], C[2 * D * Q2 + c], a1);    // d/dz part -> B_x^T
         }
         else { a0 = fma(sB[qy + Q * dy], C[c], a0); }
      }
      A[o] = a0;
      if (diff) { A[D2 * Q + o] = a1; }
   }
   __syncthreads();
   // x^T + scatter
   for (int i = t; i < D3; i += nt)
   {
      const int dx = i % D, r = i / D;
      double v = 0.0;
      for (int qx = 0; qx < Q; qx++)
      {
         const int a = qx + Q * r;
         if (diff)
         {
            v = fma(sG[qx + Q * dx], A[a], v);
            v = fma(sB[qx + Q * dx], A[D2 * Q + a], v);
         }
         else { v = fma(sB[qx + Q * dx], A[a], v); }
      }
      if (idx) { const int g = ix[i]; if (g >= 0) { atomicAdd(y + g * vxs + comp * vxo, v); } }
      else { y[eoff + i] += v; }
   }
}

// ---- 2D -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pa_apply_generic_2d_kernel(int kind, int D, int Q, int ncomp, const double *__restrict__ Bg,
                           const double *__restrict__ Gg, const double *__restrict__ qdata,
                           const double *__restrict__ x, double *__restrict__ y,
                           const int *__restrict__ idx, int vd, long long vxs, long long vxo, long long vics)
{
   // vector spaces (vdim > 1): blockIdx.y = component; component c of L-dof g is x[g*vxs + c*vxo],
   // its gather index starts at idx + c*vics, its E-vector block is (e*vd + c)
   const int comp = blockIdx.y;
   extern __shared__ double sm[];
   const bool diff = kind == B2F_KIND_DIFFUSION;
   const int D2 = D * D, Q2 = Q * Q;
   double *sB = sm, *sG = sB + Q * D;
   double *X = sG + Q * D;       // [dy][dx]
   double *A = X + D2;           // [f][dy][qx]
   double *E = A + 2 * D * Q;    // [f][qy][qx]
   const int64_t e = blockIdx.x;
   const int t = threadIdx.x, nt = blockDim.x;
   for (int i = t; i < Q * D; i += nt) { sB[i] = Bg[i]; sG[i] = Gg ? Gg[i] : 0.0; }
   const int *ix = idx ? idx + comp * vics + e * D2 : nullptr;
   const int64_t eoff = (e * vd + comp) * (int64_t)D2;
   for (int i = t; i < D2; i += nt)
   {
      if (idx) { const int g = ix[i]; X[i] = g >= 0 ? x[g * vxs + comp * vxo] : 0.0; }
      else { X[i] = x[eoff + i]; }
   }
   __syncthreads();
   for (int o = t; o < D * Q; o += nt)
   {
      const int qx = o % Q, dy = o / Q;
      double vb = 0.0, vg = 0.0;
      for (int dx = 0; dx < D; dx++)
      {
         vb = fma(sB[qx + Q * dx], X[dx + D * dy], vb);
         vg = fma(sG[qx + Q * dx], X[dx + D * dy], vg);
      }
      A[o] = vb; A[D * Q + o] = vg;
   }
   __syncthreads();
   const double *De = qdata + e * (int64_t)ncomp * Q2;
   for (int o = t; o < Q2; o += nt)
   {
      const int qx = o % Q, qy = o / Q;
      double u[2] = {0.0, 0.0}, v[2];
      for (int dy = 0; dy < D; dy++)
      {
         const int a = qx + Q * dy;
         if (diff)
         {
            u[0] = fma(sB[qy + Q * dy], A[D * Q + a], u[0]); // d/dx
            u[1] = fma(sG[qy + Q * dy], A[a], u[1]);         // d/dy
         }
         else { u[0] = fma(sB[qy + Q * dy], A[a], u[0]); }
      }
      if (diff) { apply_qdata<2>(De + o, Q2, ncomp, u, v); E[o] = v[0]; E[Q2 + o] = v[1]; }
      else { E[o] = De[o] * u[0]; }
   }
   __syncthreads();
   for (int o = t; o < D * Q; o += nt)
   {
      const int qx = o % Q, dy = o / Q;
      double a0 = 0.0, a1 = 0.0;
      for (int qy = 0; qy < Q; qy++)
      {
         const int q = qx + Q * qy;
         if (diff)
         {
            a0 = fma(sB[qy + Q * dy], E[q], a0);      // -> G_x^T
            a1 = fma(sG[qy + Q * dy], E[Q2 + q], a1); // -> B_x^T
         }
         else { a0 = fma(sB[qy + Q * dy], E[q], a0); }
      }
      A[o] = a0; A[D * Q + o] = a1;
   }
   __syncthreads();
   for (int i = t; i < D2; i += nt)
   {
      const int dx = i % D, dy = i / D;
      double v = 0.0;
      for (int qx = 0; qx < Q; qx++)
      {
         const int a = qx + Q * dy;
         if (diff)
         {
            v = fma(sG[qx + Q * dx], A[a], v);
            v = fma(sB[qx + Q * dx], A[D * Q + a], v);
         }
         else { v = fma(sB[qx + Q * dx], A[a], v); }
      }
      if (idx) { const int g = ix[i]; if (g >= 0) { atomicAdd(y + g * vxs + comp * vxo, v); } }
      else { y[eoff + i] += v; }
   }
}

int pa_apply_generic(b2f_pa *pa, const double *x, double *y, const int *idx)
{
   b2f_ctx *ctx = pa->ctx;
   const int D = pa->d1d, Q = pa->q1d;
   if (pa->ne == 0) { return B2F_OK; }
   if (pa->ne > 0x7fffffffLL)
   {
      set_error("pa_apply_generic: %lld elements exceed the grid limit", (long long)pa->ne);
      return B2F_ERR_UNSUPPORTED;
   }
   const bool diff = pa->kind == B2F_KIND_DIFFUSION;
   size_t smem;
   if (pa->dim == 3)
   {
      const int nf1 = diff ? 2 : 1, nf2 = diff ? 3 : 1;
      smem = sizeof(double) * (size_t)(2 * Q * D + D * D * D + nf1 * D * D * Q + nf2 * D * Q * Q + nf2 * Q * Q * Q);
   }
   else { smem = sizeof(double) * (size_t)(2 * Q * D + D * D + 2 * D * Q + 2 * Q * Q); }
   if ((int)smem > ctx->max_smem_optin)
   {
      set_error("pa_apply_generic: D1D=%d Q1D=%d needs %zu B of shared memory", D, Q, smem);
      return B2F_ERR_UNSUPPORTED;
   }
   const double *G = diff ? pa->d_G : nullptr;
   if (pa->dim == 3)
   {
      B2F_CUDA(cudaFuncSetAttribute(pa_apply_generic_3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      pa_apply_generic_3d_kernel<<<dim3((unsigned)pa->ne, (unsigned)pa->vd), 128, smem, ctx->stream>>>(
         pa->kind, D, Q, pa->ncomp, pa->d_B, G, pa->data, x, y, idx, pa->vd, pa->vxs, pa->vxo, pa->vics);
   }
   else
   {
      B2F_CUDA(cudaFuncSetAttribute(pa_apply_generic_2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      pa_apply_generic_2d_kernel<<<dim3((unsigned)pa->ne, (unsigned)pa->vd), 128, smem, ctx->stream>>>(
         pa->kind, D, Q, pa->ncomp, pa->d_B, G, pa->data, x, y, idx, pa->vd, pa->vxs, pa->vxo, pa->vics);
   }
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

} // namespace b2f
