// mfem_b200/csrc/pa_diag.cu -- AssembleDiagonalPA kernels (setup path, run once per solve).
//
// Replaces (Smem)PADiffusionDiagonal2D/3D (fem/integ/bilininteg_diffusion_kernels.hpp:
// 92-485) and (Smem)PAMassAssembleDiagonal2D/3D (fem/integ/bilininteg_mass_kernels.hpp:
// 132-416): diag_E(d) += sum_q  (B|G)(q,d)^2-products * D(q), sum-factorised.
//
// One CTA per element, runtime (D1D,Q1D); the intermediates QQD / QDD live in shared
// memory; each output dof is owned by one thread, so the accumulation over the (i,j)
// passes needs no atomics and is deterministic.
#include "common.cuh"

namespace b2f
{

__global__ void __launch_bounds__(128)
diffusion_diag_3d_kernel(int D1D, int Q1D, int ncomp, const double *__restrict__ Bg,
                         const double *__restrict__ Gg, const double *__restrict__ Dq,
                         double *__restrict__ Y)
{
   extern __shared__ double sm[];
   double *sB = sm, *sG = sB + Q1D * D1D;
   double *QQD = sG + Q1D * D1D;          // [qx][qy][dz]
   double *QDD = QQD + Q1D * Q1D * D1D;   // [qx][dy][dz]
   const int64_t e = blockIdx.x;
   const int Q3 = Q1D * Q1D * Q1D, D3 = D1D * D1D * D1D;
   const double *De = Dq + e * (int64_t)(ncomp * Q3);
   double *Ye = Y + e * (int64_t)D3;
   for (int i = threadIdx.x; i < Q1D * D1D; i += blockDim.x) { sB[i] = Bg[i]; sG[i] = Gg[i]; }
   __syncthreads();
   for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
      {
         // symmetric: 6 components; full matrix: row-major i*3+j (kernels.hpp:300-318)
         const int k = ncomp == 9 ? i * 3 + j
                       : j >= i ? 3 - (3 - i) * (2 - i) / 2 + j : 3 - (3 - j) * (2 - j) / 2 + i;
         for (int t = threadIdx.x; t < Q1D * Q1D * D1D; t += blockDim.x)
         {
            const int dz = t % D1D, qy = (t / D1D) % Q1D, qx = t / (D1D * Q1D);
            double s = 0.0;
            for (int qz = 0; qz < Q1D; ++qz)
            {
               const double O = De[qx + Q1D * (qy + Q1D * qz) + Q3 * k];
               const double Bz = sB[qz + Q1D * dz], Gz = sG[qz + Q1D * dz];
               const double L = i == 2 ? Gz : Bz, R = j == 2 ? Gz : Bz;
               s += L * O * R;
            }
            QQD[(qx * Q1D + qy) * D1D + dz] = s;
         }
         __syncthreads();
         for (int t = threadIdx.x; t < Q1D * D1D * D1D; t += blockDim.x)
         {
            const int dz = t % D1D, dy = (t / D1D) % D1D, qx = t / (D1D * D1D);
            double s = 0.0;
            for (int qy = 0; qy < Q1D; ++qy)
            {
               const double By = sB[qy + Q1D * dy], Gy = sG[qy + Q1D * dy];
               const double L = i == 1 ? Gy : By, R = j == 1 ? Gy : By;
               s += L * QQD[(qx * Q1D + qy) * D1D + dz] * R;
            }
            QDD[(qx * D1D + dy) * D1D + dz] = s;
         }
         __syncthreads();
         for (int t = threadIdx.x; t < D3; t += blockDim.x)
         {
            const int dx = t % D1D, dy = (t / D1D) % D1D, dz = t / (D1D * D1D);
            double s = Ye[t];
            for (int qx = 0; qx < Q1D; ++qx)
            {
               const double Bx = sB[qx + Q1D * dx], Gx = sG[qx + Q1D * dx];
               const double L = i == 0 ? Gx : Bx, R = j == 0 ? Gx : Bx;
               s += L * QDD[(qx * D1D + dy) * D1D + dz] * R;
            }
            Ye[t] = s;
         }
         __syncthreads();
      }
}

__global__ void __launch_bounds__(128)
mass_diag_3d_kernel(int D1D, int Q1D, const double *__restrict__ Bg,
                    const double *__restrict__ Dq, double *__restrict__ Y)
{
   extern __shared__ double sm[];
   double *sB = sm;
   double *QQD = sB + Q1D * D1D;
   double *QDD = QQD + Q1D * Q1D * D1D;
   const int64_t e = blockIdx.x;
   const int Q3 = Q1D * Q1D * Q1D, D3 = D1D * D1D * D1D;
   const double *De = Dq + e * (int64_t)Q3;
   double *Ye = Y + e * (int64_t)D3;
   for (int i = threadIdx.x; i < Q1D * D1D; i += blockDim.x) { sB[i] = Bg[i]; }
   __syncthreads();
   for (int t = threadIdx.x; t < Q1D * Q1D * D1D; t += blockDim.x)
   {
      const int dz = t % D1D, qy = (t / D1D) % Q1D, qx = t / (D1D * Q1D);
      double s = 0.0;
      for (int qz = 0; qz < Q1D; ++qz)
      {
         const double b = sB[qz + Q1D * dz];
         s += b * b * De[qx + Q1D * (qy + Q1D * qz)];
      }
      QQD[(qx * Q1D + qy) * D1D + dz] = s;
   }
   __syncthreads();
   for (int t = threadIdx.x; t < Q1D * D1D * D1D; t += blockDim.x)
   {
      const int dz = t % D1D, dy = (t / D1D) % D1D, qx = t / (D1D * D1D);
      double s = 0.0;
      for (int qy = 0; qy < Q1D; ++qy)
      {
         const double b = sB[qy + Q1D * dy];
         s += b * b * QQD[(qx * Q1D + qy) * D1D + dz];
      }
      QDD[(qx * D1D + dy) * D1D + dz] = s;
   }
   __syncthreads();
   for (int t = threadIdx.x; t < D3; t += blockDim.x)
   {
      const int dx = t % D1D, dy = (t / D1D) % D1D, dz = t / (D1D * D1D);
      double s = 0.0;
      for (int qx = 0; qx < Q1D; ++qx)
      {
         const double b = sB[qx + Q1D * dx];
         s += b * b * QDD[(qx * D1D + dy) * D1D + dz];
      }
      Ye[t] += s;
   }
}

__global__ void __launch_bounds__(128)
diag_2d_kernel(int kind, int D1D, int Q1D, int ncomp_in, const double *__restrict__ Bg,
               const double *__restrict__ Gg, const double *__restrict__ Dq,
               double *__restrict__ Y)
{
   extern __shared__ double sm[];
   double *sB = sm, *sG = sB + Q1D * D1D;
   double *QD0 = sG + Q1D * D1D, *QD1 = QD0 + Q1D * D1D, *QD2 = QD1 + Q1D * D1D; // [qx][dy]
   const int64_t e = blockIdx.x;
   const int Q2 = Q1D * Q1D, D2 = D1D * D1D;
   const int ncomp = kind == B2F_KIND_MASS ? 1 : ncomp_in;
   const double *De = Dq + e * (int64_t)(ncomp * Q2);
   double *Ye = Y + e * (int64_t)D2;
   for (int i = threadIdx.x; i < Q1D * D1D; i += blockDim.x) { sB[i] = Bg[i]; sG[i] = Gg[i]; }
   __syncthreads();
   for (int t = threadIdx.x; t < Q1D * D1D; t += blockDim.x)
   {
      const int dy = t % D1D, qx = t / D1D;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int qy = 0; qy < Q1D; ++qy)
      {
         const int q = qx + qy * Q1D;
         const double b = sB[qy + Q1D * dy], g = sG[qy + Q1D * dy];
         if (kind == B2F_KIND_MASS) { s0 += b * b * De[q]; }
         else
         {
            // (11, 21, 12, 22) for a full matrix, (11, 21, 22) when symmetric (kernels.hpp:130-140)
            const double D00 = De[q], D10 = De[q + Q2];
            const double D01 = ncomp == 4 ? De[q + 2 * Q2] : D10;
            const double D11 = ncomp == 4 ? De[q + 3 * Q2] : De[q + 2 * Q2];
            s0 += b * b * D00;
            s1 += b * g * (D01 + D10);
            s2 += g * g * D11;
         }
      }
      QD0[qx * D1D + dy] = s0; QD1[qx * D1D + dy] = s1; QD2[qx * D1D + dy] = s2;
   }
   __syncthreads();
   for (int t = threadIdx.x; t < D2; t += blockDim.x)
   {
      const int dx = t % D1D, dy = t / D1D;
      double s = Ye[t];
      for (int qx = 0; qx < Q1D; ++qx)
      {
         const double b = sB[qx + Q1D * dx], g = sG[qx + Q1D * dx];
         if (kind == B2F_KIND_MASS) { s += b * b * QD0[qx * D1D + dy]; }
         else
         {
            s += g * g * QD0[qx * D1D + dy];
            s += g * b * QD1[qx * D1D + dy];
            s += b * b * QD2[qx * D1D + dy];
         }
      }
      Ye[t] = s;
   }
}

int pa_diag(b2f_pa *pa, double *diag_e)
{
   b2f_ctx *ctx = pa->ctx;
   if (pa->ne == 0) { return B2F_OK; }
   const int D = pa->d1d, Q = pa->q1d;
   const unsigned grid = (unsigned)pa->ne;
   if (pa->dim == 3)
   {
      const size_t smem = sizeof(double) * (2 * Q * D + Q * Q * D + Q * D * D);
      if (pa->kind == B2F_KIND_DIFFUSION)
      {
         diffusion_diag_3d_kernel<<<grid, 128, smem, ctx->stream>>>(D, Q, pa->ncomp, pa->d_B, pa->d_G,
                                                                    pa->data, diag_e);
      }
      else
      {
         mass_diag_3d_kernel<<<grid, 128, smem, ctx->stream>>>(D, Q, pa->d_B, pa->data, diag_e);
      }
   }
   else
   {
      const size_t smem = sizeof(double) * (5 * Q * D);
      diag_2d_kernel<<<grid, 128, smem, ctx->stream>>>(pa->kind, D, Q, pa->ncomp, pa->d_B, pa->d_G,
                                                       pa->data, diag_e);
   }
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

} // namespace b2f
