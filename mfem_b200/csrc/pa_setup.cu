// mfem_b200/csrc/pa_setup.cu -- quadrature-point data ("pa_data") assembly.
//
// Replaces PADiffusionSetup2D/3D (fem/integ/bilininteg_diffusion_kernels.cpp:146-386) and
// the MassIntegrator::AssemblePA lambda (fem/integ/bilininteg_mass_pa.cpp:73-90); the
// *_from_nodes variants additionally fold in GeometricFactors (mesh/mesh.cpp:15420-15498,
// fem/qinterp/grad.hpp:235, det.hpp) so the Jacobians -- 9 Q^3 doubles per element, the
// largest allocation of a reference run -- never exist in HBM.
//
// One thread per quadrature point, q fastest: every load of J (component stride Q^dim)
// and every store of D is a fully coalesced FP64 stream.  HBM-bound, no reuse.
#include "common.cuh"

namespace b2f
{

// D = (w/detJ) adj(J) C adj(J)^T ; identical operation order to the reference so that
// D matches bit for bit (FMA contraction is disabled in this file's arithmetic via
// __dmul_rn/__dadd_rn where the reference has separate roundings).
struct Sym3
{
   double d[9];
};

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

// bilininteg_diffusion_kernels.cpp:292-381 for one point; out has 6 (symmetric) or 9 entries
__device__ __forceinline__ void diffusion_point_3d(const double J[9] /* J(i,j) = J[i+3j] */,
                                                   double w, const double *C, int coeffDim,
                                                   double *out)
{
   const bool symmetric = coeffDim != 9;
   const double J11 = J[0], J21 = J[1], J31 = J[2];
   const double J12 = J[3], J22 = J[4], J32 = J[5];
   const double J13 = J[6], J23 = J[7], J33 = J[8];
   const double detJ = add(sub(mul(J11, sub(mul(J22, J33), mul(J32, J23))),
                               mul(J21, sub(mul(J12, J33), mul(J32, J13)))),
                           mul(J31, sub(mul(J12, J23), mul(J22, J13))));
   const double w_detJ = w / detJ;
   const double A11 = sub(mul(J22, J33), mul(J23, J32));
   const double A12 = sub(mul(J32, J13), mul(J12, J33));
   const double A13 = sub(mul(J12, J23), mul(J22, J13));
   const double A21 = sub(mul(J31, J23), mul(J21, J33));
   const double A22 = sub(mul(J11, J33), mul(J13, J31));
   const double A23 = sub(mul(J21, J13), mul(J11, J23));
   const double A31 = sub(mul(J21, J32), mul(J31, J22));
   const double A32 = sub(mul(J31, J12), mul(J11, J32));
   const double A33 = sub(mul(J11, J22), mul(J12, J21));
#define S3(a, b, c) add(add(a, b), c)
   if (coeffDim == 6 || coeffDim == 9)
   {
      const double M11 = C[0], M12 = C[1], M13 = C[2];
      const double M21 = (!symmetric) ? C[3] : M12;
      const double M22 = (!symmetric) ? C[4] : C[3];
      const double M23 = (!symmetric) ? C[5] : C[4];
      const double M31 = (!symmetric) ? C[6] : M13;
      const double M32 = (!symmetric) ? C[7] : M23;
      const double M33 = (!symmetric) ? C[8] : C[5];
      const double R11 = S3(mul(M11, A11), mul(M12, A12), mul(M13, A13));
      const double R12 = S3(mul(M11, A21), mul(M12, A22), mul(M13, A23));
      const double R13 = S3(mul(M11, A31), mul(M12, A32), mul(M13, A33));
      const double R21 = S3(mul(M21, A11), mul(M22, A12), mul(M23, A13));
      const double R22 = S3(mul(M21, A21), mul(M22, A22), mul(M23, A23));
      const double R23 = S3(mul(M21, A31), mul(M22, A32), mul(M23, A33));
      const double R31 = S3(mul(M31, A11), mul(M32, A12), mul(M33, A13));
      const double R32 = S3(mul(M31, A21), mul(M32, A22), mul(M33, A23));
      const double R33 = S3(mul(M31, A31), mul(M32, A32), mul(M33, A33));
      out[0] = mul(w_detJ, S3(mul(A11, R11), mul(A12, R21), mul(A13, R31)));
      out[1] = mul(w_detJ, S3(mul(A11, R12), mul(A12, R22), mul(A13, R32)));
      out[2] = mul(w_detJ, S3(mul(A11, R13), mul(A12, R23), mul(A13, R33)));
      const double D22 = mul(w_detJ, S3(mul(A21, R12), mul(A22, R22), mul(A23, R32)));
      const double D23 = mul(w_detJ, S3(mul(A21, R13), mul(A22, R23), mul(A23, R33)));
      const double D33 = mul(w_detJ, S3(mul(A31, R13), mul(A32, R23), mul(A33, R33)));
      out[4] = symmetric ? D23 : D22;
      out[5] = symmetric ? D33 : D23;
      if (symmetric) { out[3] = D22; }
      else
      {
         out[3] = mul(w_detJ, S3(mul(A21, R11), mul(A22, R21), mul(A23, R31)));
         out[6] = mul(w_detJ, S3(mul(A31, R11), mul(A32, R21), mul(A33, R31)));
         out[7] = mul(w_detJ, S3(mul(A31, R12), mul(A32, R22), mul(A33, R32)));
         out[8] = D33;
      }
   }
   else
   {
      const double C1 = C[0];
      const double C2 = C[coeffDim == 3 ? 1 : 0];
      const double C3 = C[coeffDim == 3 ? 2 : 0];
      out[0] = mul(w_detJ, S3(mul(mul(C1, A11), A11), mul(mul(C2, A12), A12), mul(mul(C3, A13), A13)));
      out[1] = mul(w_detJ, S3(mul(mul(C1, A11), A21), mul(mul(C2, A12), A22), mul(mul(C3, A13), A23)));
      out[2] = mul(w_detJ, S3(mul(mul(C1, A11), A31), mul(mul(C2, A12), A32), mul(mul(C3, A13), A33)));
      out[3] = mul(w_detJ, S3(mul(mul(C1, A21), A21), mul(mul(C2, A22), A22), mul(mul(C3, A23), A23)));
      out[4] = mul(w_detJ, S3(mul(mul(C1, A21), A31), mul(mul(C2, A22), A32), mul(mul(C3, A23), A33)));
      out[5] = mul(w_detJ, S3(mul(mul(C1, A31), A31), mul(mul(C2, A32), A32), mul(mul(C3, A33), A33)));
   }
#undef S3
}

// bilininteg_diffusion_kernels.cpp:175-211 for one point; out has 3 or 4 entries
__device__ __forceinline__ void diffusion_point_2d(const double J[4] /* J(i,j) = J[i+2j] */,
                                                   double w, const double *C, int coeffDim,
                                                   double *out)
{
   const bool symmetric = coeffDim != 4;
   const double J11 = J[0], J21 = J[1], J12 = J[2], J22 = J[3];
   const double w_detJ = w / sub(mul(J11, J22), mul(J21, J12));
   if (coeffDim == 3 || coeffDim == 4)
   {
      const double M11 = C[0], M12 = C[1];
      const double M21 = symmetric ? M12 : C[2];
      const double M22 = symmetric ? C[2] : C[3];
      const double R11 = sub(mul(M11, J22), mul(M12, J12));
      const double R21 = sub(mul(M21, J22), mul(M22, J12));
      const double R12 = add(mul(-M11, J21), mul(M12, J11));
      const double R22 = add(mul(-M21, J21), mul(M22, J11));
      out[0] = mul(w_detJ, sub(mul(J22, R11), mul(J12, R21)));
      out[1] = mul(w_detJ, add(mul(-J21, R11), mul(J11, R21)));
      out[2] = mul(w_detJ, symmetric ? add(mul(-J21, R12), mul(J11, R22))
                   : sub(mul(J22, R12), mul(J12, R22)));
      if (!symmetric) { out[3] = mul(w_detJ, add(mul(-J21, R12), mul(J11, R22))); }
   }
   else
   {
      const double C1 = C[0];
      const double C2 = C[coeffDim == 2 ? 1 : 0];
      out[0] = mul(w_detJ, add(mul(mul(C2, J12), J12), mul(mul(C1, J22), J22)));
      out[1] = mul(-w_detJ, add(mul(mul(C2, J12), J11), mul(mul(C1, J22), J21)));
      out[2] = mul(w_detJ, add(mul(mul(C2, J11), J11), mul(mul(C1, J21), J21)));
   }
}

template <int DIM>
__global__ void __launch_bounds__(256)
diffusion_setup_kernel(int NQ, int64_t ne, const double *__restrict__ W,
                       const double *__restrict__ J, const double *__restrict__ C,
                       int coeffDim, int const_c, double *__restrict__ D)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= ne * NQ) { return; }
   const int64_t e = i / NQ;
   const int q = (int)(i - e * NQ);
   constexpr int JJ = DIM * DIM;
   const int ncomp = (coeffDim == JJ) ? JJ : (DIM * (DIM + 1)) / 2;
   double Jl[JJ], Cl[JJ], out[JJ];
#pragma unroll
   for (int k = 0; k < JJ; k++) { Jl[k] = __ldg(J + q + (int64_t)NQ * (k + JJ * e)); }
   for (int k = 0; k < coeffDim; k++)
   {
      Cl[k] = const_c ? __ldg(C + k) : __ldg(C + k + (int64_t)coeffDim * (q + (int64_t)NQ * e));
   }
   if (DIM == 3) { diffusion_point_3d(Jl, __ldg(W + q), Cl, coeffDim, out); }
   else { diffusion_point_2d(Jl, __ldg(W + q), Cl, coeffDim, out); }
   for (int k = 0; k < ncomp; k++) { D[q + (int64_t)NQ * (k + (int64_t)ncomp * e)] = out[k]; }
}

int diffusion_setup(b2f_ctx *ctx, int dim, int q1d, int64_t ne, const double *W_d,
                    const double *J, const double *coeff, int coeff_dim, int coeff_const,
                    double *D)
{
   const int NQ = dim == 3 ? q1d * q1d * q1d : q1d * q1d;
   const int64_t n = ne * NQ;
   if (n == 0) { return B2F_OK; }
   const unsigned grid = (unsigned)cdiv(n, 256);
   if (dim == 3)
   {
      diffusion_setup_kernel<3><<<grid, 256, 0, ctx->stream>>>(NQ, ne, W_d, J, coeff, coeff_dim,
                                                              coeff_const, D);
   }
   else
   {
      diffusion_setup_kernel<2><<<grid, 256, 0, ctx->stream>>>(NQ, ne, W_d, J, coeff, coeff_dim,
                                                              coeff_const, D);
   }
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

// bilininteg_mass_pa.cpp:84-89
__global__ void __launch_bounds__(256)
mass_setup_kernel(int NQ, int64_t n, const double *__restrict__ W, const double *__restrict__ detJ,
                  const double *__restrict__ C, int const_c, double *__restrict__ D)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   const int q = (int)(i % NQ);
   const double c = const_c ? __ldg(C) : __ldg(C + i);
   D[i] = mul(mul(__ldg(W + q), c), __ldg(detJ + i));
}

int mass_setup(b2f_ctx *ctx, int nq, int64_t ne, const double *W_d, const double *detJ,
               const double *coeff, int coeff_const, double *D)
{
   const int64_t n = ne * nq;
   if (n == 0) { return B2F_OK; }
   mass_setup_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(nq, n, W_d, detJ, coeff,
                                                                      coeff_const, D);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

// Fused GeometricFactors + setup: J(q) = sum_nodes X_i(node) dphi_node/dxi_j(q), evaluated
// per quadrature point from the element's nodal coordinates (L1/L2 resident: (nd^dim, dim)
// doubles per element).  Bn, Gn: (Q1D, nd) tables of the nodal basis.
template <int DIM>
__global__ void __launch_bounds__(256)
setup_from_nodes_kernel(int kind, int q1d, int nd, int64_t ne, const double *__restrict__ W,
                        const double *__restrict__ Bn, const double *__restrict__ Gn,
                        const double *__restrict__ X, double coeff, double *__restrict__ D)
{
   extern __shared__ double sbg[]; // Bn | Gn
   const int ntab = q1d * nd;
   for (int i = threadIdx.x; i < ntab; i += blockDim.x) { sbg[i] = Bn[i]; sbg[ntab + i] = Gn[i]; }
   __syncthreads();
   const double *sB = sbg, *sG = sbg + ntab;
   const int NQ = DIM == 3 ? q1d * q1d * q1d : q1d * q1d;
   const int ND = DIM == 3 ? nd * nd * nd : nd * nd;
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= ne * NQ) { return; }
   const int64_t e = i / NQ;
   const int q = (int)(i - e * NQ);
   const int qx = q % q1d, qy = (q / q1d) % q1d, qz = q / (q1d * q1d);
   const double *Xe = X + e * (int64_t)(ND * DIM);
   double Jl[DIM * DIM];
#pragma unroll
   for (int k = 0; k < DIM * DIM; k++) { Jl[k] = 0.0; }
   if (DIM == 3)
   {
      for (int dz = 0; dz < nd; dz++)
      {
         const double bz = sB[qz + q1d * dz], gz = sG[qz + q1d * dz];
         for (int dy = 0; dy < nd; dy++)
         {
            const double by = sB[qy + q1d * dy], gy = sG[qy + q1d * dy];
            double sb[3] = {0, 0, 0}, sg[3] = {0, 0, 0};
            for (int dx = 0; dx < nd; dx++)
            {
               const double bx = sB[qx + q1d * dx], gx = sG[qx + q1d * dx];
               const int n = dx + nd * (dy + nd * dz);
#pragma unroll
               for (int c = 0; c < 3; c++)
               {
                  const double v = __ldg(Xe + n + ND * c);
                  sb[c] = fma(bx, v, sb[c]);
                  sg[c] = fma(gx, v, sg[c]);
               }
            }
#pragma unroll
            for (int c = 0; c < 3; c++)
            {
               Jl[c + 3 * 0] = fma(sg[c], by * bz, Jl[c + 3 * 0]);
               Jl[c + 3 * 1] = fma(sb[c], gy * bz, Jl[c + 3 * 1]);
               Jl[c + 3 * 2] = fma(sb[c], by * gz, Jl[c + 3 * 2]);
            }
         }
      }
   }
   else
   {
      for (int dy = 0; dy < nd; dy++)
      {
         const double by = sB[qy + q1d * dy], gy = sG[qy + q1d * dy];
         double sb[2] = {0, 0}, sg[2] = {0, 0};
         for (int dx = 0; dx < nd; dx++)
         {
            const double bx = sB[qx + q1d * dx], gx = sG[qx + q1d * dx];
#pragma unroll
            for (int c = 0; c < 2; c++)
            {
               const double v = __ldg(Xe + dx + nd * dy + ND * c);
               sb[c] = fma(bx, v, sb[c]);
               sg[c] = fma(gx, v, sg[c]);
            }
         }
#pragma unroll
         for (int c = 0; c < 2; c++)
         {
            Jl[c + 2 * 0] = fma(sg[c], by, Jl[c + 2 * 0]);
            Jl[c + 2 * 1] = fma(sb[c], gy, Jl[c + 2 * 1]);
         }
      }
   }
   const double w = __ldg(W + q);
   if (kind == B2F_KIND_MASS)
   {
      double det;
      if (DIM == 3)
      {
         det = Jl[0] * (Jl[4] * Jl[8] - Jl[5] * Jl[7]) - Jl[1] * (Jl[3] * Jl[8] - Jl[5] * Jl[6]) +
               Jl[2] * (Jl[3] * Jl[7] - Jl[4] * Jl[6]);
      }
      else { det = Jl[0] * Jl[3] - Jl[1] * Jl[2]; }
      D[i] = mul(mul(w, coeff), det);
   }
   else
   {
      constexpr int NC = (DIM * (DIM + 1)) / 2;
      double out[DIM * DIM];
      if (DIM == 3) { diffusion_point_3d(Jl, w, &coeff, 1, out); }
      else { diffusion_point_2d(Jl, w, &coeff, 1, out); }
#pragma unroll
      for (int k = 0; k < NC; k++) { D[q + (int64_t)NQ * (k + (int64_t)NC * e)] = out[k]; }
   }
}

int setup_from_nodes(b2f_ctx *ctx, int kind, int dim, int q1d, int64_t ne, const double *W_d,
                     int nd, const double *Bn_d, const double *Gn_d, const double *nodes_e,
                     double coeff, double *D)
{
   const int NQ = dim == 3 ? q1d * q1d * q1d : q1d * q1d;
   const int64_t n = ne * NQ;
   if (n == 0) { return B2F_OK; }
   const size_t smem = sizeof(double) * 2 * q1d * nd;
   if (smem > 48 * 1024)
   {
      if ((int)smem > ctx->max_smem_optin)
      {
         set_error("setup_from_nodes: Q1D=%d node_d1d=%d needs %zu B of shared memory", q1d, nd, smem);
         return B2F_ERR_UNSUPPORTED;
      }
      B2F_CUDA(cudaFuncSetAttribute(setup_from_nodes_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      B2F_CUDA(cudaFuncSetAttribute(setup_from_nodes_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   }
   const unsigned grid = (unsigned)cdiv(n, 256);
   if (dim == 3)
   {
      setup_from_nodes_kernel<3><<<grid, 256, smem, ctx->stream>>>(kind, q1d, nd, ne, W_d, Bn_d,
                                                                   Gn_d, nodes_e, coeff, D);
   }
   else
   {
      setup_from_nodes_kernel<2><<<grid, 256, smem, ctx->stream>>>(kind, q1d, nd, ne, W_d, Bn_d,
                                                                   Gn_d, nodes_e, coeff, D);
   }
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

} // namespace b2f
