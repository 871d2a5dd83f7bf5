// lf_domain.cu -- DomainLFIntegrator / VectorDomainLFIntegrator fast (device) assembly:
// b_e(d, c, e) += sum_q B(q, d) W(q) f_c(q, e) detJ(q, e)    (tensor-product H1 elements)
//
// Reference: DLFEvalAssemble (fem/integ/lininteg_domain.cpp:21-47) -> DLFEvalAssemble2D/3D
// (fem/integ/lininteg_domain_kernels.hpp:76-161,165-297), driven by
// LinearFormExtension::Assemble (fem/linearform_ext.cpp:20-66).  Setup-time work (once per
// right-hand side): one CTA per element, run-time (D1D, Q1D), the contraction order of the
// reference (z, y, x in 3D; x, y in 2D).
#include "common.cuh"

namespace b2f
{

// 3D: shared cube s[(a*n + b)*n + c], n = max(d, q)
__global__ void lf_domain_3d_kernel(int d, int q, int vdim, int64_t ne, const int *__restrict__ markers,
                                    const double *__restrict__ Bg, const double *__restrict__ W,
                                    const double *__restrict__ detJ, const double *__restrict__ coeff,
                                    int cst, double *__restrict__ y)
{
   extern __shared__ double sm[];
   const int n = d > q ? d : q;
   double *Bt = sm;             // Bt[dd + d*qq] = B(qq, dd)
   double *s = sm + d * q;      // n^3
   const int64_t e = blockIdx.x;
   if (markers && markers[e] == 0) { return; }
   const int t = threadIdx.x, nt = blockDim.x;
   const int q2 = q * q, q3 = q2 * q, d3 = d * d * d;
   for (int i = t; i < d * q; i += nt) { Bt[(i / q) + d * (i % q)] = Bg[i]; } // Bg is (q, d), q fastest
   for (int c = 0; c < vdim; c++)
   {
      for (int i = t; i < q3; i += nt)
      {
         const int x = i % q, yy = (i / q) % q, z = i / q2;
         const double cv = cst ? coeff[c] : coeff[c + vdim * (i + q3 * e)];
         s[(z * n + yy) * n + x] = W[i] * cv * detJ[i + q3 * e];
      }
      __syncthreads();
      // z: (qz, qy, qx) -> (dz, qy, qx), in place per (qy, qx) column
      for (int i = t; i < q2; i += nt)
      {
         const int qx = i % q, qy = i / q;
         double u[16];
         for (int dz = 0; dz < d; dz++) { u[dz] = 0.0; }
         for (int qz = 0; qz < q; qz++)
         {
            const double v = s[(qz * n + qy) * n + qx];
            for (int dz = 0; dz < d; dz++) { u[dz] += v * Bt[dz + d * qz]; }
         }
         for (int dz = 0; dz < d; dz++) { s[(dz * n + qy) * n + qx] = u[dz]; }
      }
      __syncthreads();
      // y: (dz, qy, qx) -> (dz, dy, qx)
      for (int i = t; i < d * q; i += nt)
      {
         const int qx = i % q, dz = i / q;
         double u[16];
         for (int dy = 0; dy < d; dy++) { u[dy] = 0.0; }
         for (int qy = 0; qy < q; qy++)
         {
            const double v = s[(dz * n + qy) * n + qx];
            for (int dy = 0; dy < d; dy++) { u[dy] += v * Bt[dy + d * qy]; }
         }
         for (int dy = 0; dy < d; dy++) { s[(dz * n + dy) * n + qx] = u[dy]; }
      }
      __syncthreads();
      // x: (dz, dy, qx) -> Y(dx, dy, dz, c, e)
      for (int i = t; i < d * d; i += nt)
      {
         const int dy = i % d, dz = i / d;
         double u[16];
         for (int dx = 0; dx < d; dx++) { u[dx] = 0.0; }
         for (int qx = 0; qx < q; qx++)
         {
            const double v = s[(dz * n + dy) * n + qx];
            for (int dx = 0; dx < d; dx++) { u[dx] += v * Bt[dx + d * qx]; }
         }
         double *ye = y + ((int64_t)e * vdim + c) * d3 + (dz * d + dy) * d;
         for (int dx = 0; dx < d; dx++) { ye[dx] += u[dx]; }
      }
      __syncthreads();
   }
}

__global__ void lf_domain_2d_kernel(int d, int q, int vdim, int64_t ne, const int *__restrict__ markers,
                                    const double *__restrict__ Bg, const double *__restrict__ W,
                                    const double *__restrict__ detJ, const double *__restrict__ coeff,
                                    int cst, double *__restrict__ y)
{
   extern __shared__ double sm[];
   double *Bt = sm;            // Bt[dd + d*qq]
   double *QQ = Bt + d * q;    // QQ[qy + q*qx]
   double *QD = QQ + q * q;    // QD[qy + q*dx]
   const int64_t e = blockIdx.x;
   if (markers && markers[e] == 0) { return; }
   const int t = threadIdx.x, nt = blockDim.x;
   const int q2 = q * q, d2 = d * d;
   for (int i = t; i < d * q; i += nt) { Bt[(i / q) + d * (i % q)] = Bg[i]; }
   for (int c = 0; c < vdim; c++)
   {
      for (int i = t; i < q2; i += nt)
      {
         const int x = i % q, yy = i / q;
         const double cv = cst ? coeff[c] : coeff[c + vdim * (i + q2 * e)];
         QQ[yy + q * x] = W[i] * cv * detJ[i + q2 * e];
      }
      __syncthreads();
      for (int i = t; i < q * d; i += nt)
      {
         const int dx = i % d, qy = i / d;
         double u = 0.0;
         for (int qx = 0; qx < q; qx++) { u += QQ[qy + q * qx] * Bt[dx + d * qx]; }
         QD[qy + q * dx] = u;
      }
      __syncthreads();
      for (int i = t; i < d2; i += nt)
      {
         const int dx = i % d, dy = i / d;
         double u = 0.0;
         for (int qy = 0; qy < q; qy++) { u += QD[qy + q * dx] * Bt[dy + d * qy]; }
         y[((int64_t)e * vdim + c) * d2 + dy * d + dx] += u;
      }
      __syncthreads();
   }
}

int lf_domain(b2f_ctx *ctx, int dim, int d, int q, int vdim, int64_t ne, const int *markers,
              const double *B_d, const double *W_d, const double *detJ, const double *coeff, int cst,
              double *y_e)
{
   if (ne == 0) { return B2F_OK; }
   if (dim == 3)
   {
      const int n = d > q ? d : q;
      const size_t smem = sizeof(double) * ((size_t)d * q + (size_t)n * n * n);
      if (smem > 48 * 1024)
      {
         if ((int)smem > ctx->max_smem_optin)
         {
            set_error("lf_domain: D1D=%d Q1D=%d needs %zu B of shared memory", d, q, smem);
            return B2F_ERR_UNSUPPORTED;
         }
         B2F_CUDA(cudaFuncSetAttribute(lf_domain_3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
      }
      const int nt = ((q * q + 31) / 32) * 32;
      lf_domain_3d_kernel<<<(unsigned)ne, nt, smem, ctx->stream>>>(d, q, vdim, ne, markers, B_d, W_d, detJ,
                                                                   coeff, cst, y_e);
   }
   else
   {
      const size_t smem = sizeof(double) * ((size_t)2 * d * q + (size_t)q * q);
      if (smem > 48 * 1024)
      {
         if ((int)smem > ctx->max_smem_optin)
         {
            set_error("lf_domain: D1D=%d Q1D=%d needs %zu B of shared memory", d, q, smem);
            return B2F_ERR_UNSUPPORTED;
         }
         B2F_CUDA(cudaFuncSetAttribute(lf_domain_2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
      }
      const int nt = ((q * q + 31) / 32) * 32;
      lf_domain_2d_kernel<<<(unsigned)ne, nt, smem, ctx->stream>>>(d, q, vdim, ne, markers, B_d, W_d, detJ,
                                                                   coeff, cst, y_e);
   }
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

} // namespace b2f
