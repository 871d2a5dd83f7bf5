// mfem_b200/csrc/lor.cu -- low-order-refined (LOR) assembly of H1 diffusion + mass forms in 3D.
//
// What it replaces: BatchedLORAssembly::Assemble for H1 spaces (fem/lor/lor_batched.cpp:712-734):
//   FormLORVertexCoordinates (lor_batched.cpp:75-112), BatchedLOR_H1::Assemble3D
//   (fem/lor/lor_h1_impl.hpp:184-455), FillI / FillJAndData (lor_batched.cpp:143-369) and
//   SparseMatrix::EliminateBC(ess, DIAG_KEEP) (linalg/sparsemat.cpp:2375-2416).  The result is the
//   CSR matrix MFEM hands to an algebraic multigrid (LORSolver<HypreBoomerAMG>, fem/lor/lor.hpp).
//
// How it differs from the reference's device path:
//   * the macro element's 27-point stencil array lives in shared memory while its p^3 sub-hexahedra
//     are added, in 8 parity colours (sub-elements of one colour share no vertex), so there are no
//     atomics and the result is reproducible bit for bit; the reference adds with global atomics;
//   * the CSR is filled by ROWS (one thread per dof walks the elements holding it, in element order),
//     not by (element, local dof) with an atomic cursor per row: the column order inside a row is then
//     exactly the order of the reference's serial host loop, so I and J match the reference's CPU result
//     bit for bit (its own device path returns the columns of a row in a run-dependent order);
//   * the stencil index of a pair of dofs in another element is computed from their lexicographic
//     coordinates instead of searching the 27-entry map.
#include "common.cuh"

#include <cub/device/device_scan.cuh>
#include <new>

struct b2f_lor
{
   b2f_ctx *ctx = nullptr;
   int p = 0, ndofs = 0;
   int64_t ne = 0, nnz = 0;
   int *I = nullptr, *J = nullptr;
   double *A = nullptr;
   double *V = nullptr; // (27, nd1d^3, NE) macro-element stencil array ("sparse_ij")
};

namespace b2f
{

// ---- vertex coordinates -------------------------------------------------------------------------
// X_vert(c, i, e) = sum_j Bn(ix,jx) Bn(iy,jy) Bn(iz,jz) nodes(j, c, e); Bn: (p+1, nd) row-major
__global__ void lor_vertex_coordinates_kernel(int n1, int nd, int64_t ne, const double *__restrict__ Bn,
                                              const double *__restrict__ nodes, double *__restrict__ xv)
{
   const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int nv = n1 * n1 * n1, nn = nd * nd * nd;
   if (tid >= ne * nv) { return; }
   const int64_t e = tid / nv;
   const int i = (int)(tid - e * nv);
   const int ix = i % n1, iy = (i / n1) % n1, iz = i / (n1 * n1);
   double s[3] = {0.0, 0.0, 0.0};
   const double *ne_ = nodes + e * 3 * (int64_t)nn;
   for (int jz = 0; jz < nd; jz++)
   {
      const double bz = Bn[iz * nd + jz];
      for (int jy = 0; jy < nd; jy++)
      {
         const double byz = bz * Bn[iy * nd + jy];
         for (int jx = 0; jx < nd; jx++)
         {
            const double b = byz * Bn[ix * nd + jx];
            const int j = jx + nd * (jy + nd * jz);
            s[0] += b * ne_[j];
            s[1] += b * ne_[nn + j];
            s[2] += b * ne_[2 * nn + j];
         }
      }
   }
   double *o = xv + 3 * tid;
   o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
}

// ---- macro-element stencil array ----------------------------------------------------------------
// One sub-hexahedron: trilinear map, vertex quadrature (weight 1/8).  P[c][x][y][z] are its corner
// coordinates, mq / dq the coefficients at the corners.  M is the lower triangle of the symmetric
// 8x8 matrix, local dof ii = ix + 2 iy + 4 iz.
__device__ __forceinline__ void lor_local_matrix(const double (&P)[3][2][2][2], const double (&mq)[2][2][2],
                                                 const double (&dq)[2][2][2], double (&M)[8][8])
{
#pragma unroll
   for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) { M[i][j] = 0.0; }
#pragma unroll
   for (int qz = 0; qz < 2; qz++)
#pragma unroll
      for (int qy = 0; qy < 2; qy++)
#pragma unroll
         for (int qx = 0; qx < 2; qx++)
         {
            // at a corner the Jacobian of the trilinear map is the three edge differences
            double J[3][3];
#pragma unroll
            for (int c = 0; c < 3; c++)
            {
               J[c][0] = P[c][1][qy][qz] - P[c][0][qy][qz];
               J[c][1] = P[c][qx][1][qz] - P[c][qx][0][qz];
               J[c][2] = P[c][qx][qy][1] - P[c][qx][qy][0];
            }
            double A[3][3];
            A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
            A[0][1] = J[2][1] * J[0][2] - J[0][1] * J[2][2];
            A[0][2] = J[0][1] * J[1][2] - J[1][1] * J[0][2];
            A[1][0] = J[2][0] * J[1][2] - J[1][0] * J[2][2];
            A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
            A[1][2] = J[1][0] * J[0][2] - J[0][0] * J[1][2];
            A[2][0] = J[1][0] * J[2][1] - J[2][0] * J[1][1];
            A[2][1] = J[2][0] * J[0][1] - J[0][0] * J[2][1];
            A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
            const double detJ = J[0][0] * A[0][0] + J[1][0] * A[0][1] + J[2][0] * A[0][2];
            const double w = 0.125;
            const double wd = dq[qx][qy][qz] * (w / detJ);
            double K[3][3];
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
               for (int b = 0; b <= a; b++)
               {
                  K[a][b] = wd * (A[a][0] * A[b][0] + A[a][1] * A[b][1] + A[a][2] * A[b][2]);
                  K[b][a] = K[a][b];
               }
            // d/dxi_a of shape i at corner q is -1 / +1 for the two shapes that agree with q in the
            // other two directions, 0 for the rest
            const int q3[3] = {qx, qy, qz};
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
               for (int b = 0; b < 3; b++)
#pragma unroll
                  for (int s = 0; s < 2; s++)
#pragma unroll
                     for (int t = 0; t < 2; t++)
                     {
                        int i3[3] = {q3[0], q3[1], q3[2]}, j3[3] = {q3[0], q3[1], q3[2]};
                        i3[a] = s; j3[b] = t;
                        const int ii = i3[0] + 2 * i3[1] + 4 * i3[2], jj = j3[0] + 2 * j3[1] + 4 * j3[2];
                        if (jj <= ii) { M[ii][jj] += ((s == t) ? K[a][b] : -K[a][b]); }
                     }
            const int iq = qx + 2 * qy + 4 * qz;
            M[iq][iq] += mq[qx][qy][qz] * (w * detJ); // vertex quadrature lumps the mass matrix
         }
}

// CTA = `neb` macro elements; Vs[(el * nv + i) * 27 + j] in shared memory
__global__ void __launch_bounds__(128)
lor_h1_assemble_3d_kernel(int p, int neb, int64_t ne, const double *__restrict__ xv,
                          const double *__restrict__ cm, int cm_const, const double *__restrict__ cd,
                          int cd_const, double *__restrict__ V)
{
   extern __shared__ double Vs[];
   const int n1 = p + 1, nv = n1 * n1 * n1;
   const int64_t e0 = (int64_t)blockIdx.x * neb;
   const int nev = (int)((ne - e0) < neb ? (ne - e0) : neb);
   for (int i = threadIdx.x; i < nev * nv * 27; i += blockDim.x) { Vs[i] = 0.0; }
   __syncthreads();
   for (int colour = 0; colour < 8; colour++)
   {
      const int cx = colour & 1, cy = (colour >> 1) & 1, cz = colour >> 2;
      const int mx = (p - cx + 1) / 2, my = (p - cy + 1) / 2, mz = (p - cz + 1) / 2; // sub-elements per direction
      const int per_el = mx * my * mz;
      for (int s = threadIdx.x; s < nev * per_el; s += blockDim.x)
      {
         const int el = s / per_el;
         int r = s - el * per_el;
         const int kx = 2 * (r % mx) + cx;
         r /= mx;
         const int ky = 2 * (r % my) + cy, kz = 2 * (r / my) + cz;
         const int64_t e = e0 + el;
         double P[3][2][2][2], mq[2][2][2], dq[2][2][2];
#pragma unroll
         for (int dz = 0; dz < 2; dz++)
#pragma unroll
            for (int dy = 0; dy < 2; dy++)
#pragma unroll
               for (int dx = 0; dx < 2; dx++)
               {
                  const int64_t v = (kx + dx) + n1 * ((ky + dy) + n1 * (kz + dz)) + (int64_t)nv * e;
                  P[0][dx][dy][dz] = xv[3 * v];
                  P[1][dx][dy][dz] = xv[3 * v + 1];
                  P[2][dx][dy][dz] = xv[3 * v + 2];
                  mq[dx][dy][dz] = cm_const ? cm[0] : cm[v];
                  dq[dx][dy][dz] = cd_const ? cd[0] : cd[v];
               }
         double M[8][8];
         lor_local_matrix(P, mq, dq, M);
         double *Ve = Vs + (size_t)el * nv * 27;
#pragma unroll
         for (int ii = 0; ii < 8; ii++)
         {
            const int ix = ii & 1, iy = (ii >> 1) & 1, iz = ii >> 2;
            double *row = Ve + ((ix + kx) + n1 * ((iy + ky) + n1 * (iz + kz))) * 27;
#pragma unroll
            for (int jj = 0; jj < 8; jj++)
            {
               const int jx = jj & 1, jy = (jj >> 1) & 1, jz = jj >> 2;
               const int off = (jx - ix + 1) + 3 * (jy - iy + 1) + 9 * (jz - iz + 1);
               row[off] += (jj <= ii) ? M[ii][jj] : M[jj][ii];
            }
         }
      }
      __syncthreads();
   }
   double *Vg = V + e0 * (int64_t)nv * 27;
   for (int i = threadIdx.x; i < nev * nv * 27; i += blockDim.x) { Vg[i] = Vs[i]; }
}

// ---- CSR ------------------------------------------------------------------------------------------
__device__ __forceinline__ int lor_decode(int s) { return s >= 0 ? s : -1 - s; }

// Walk row `ii` as the reference's serial loop meets it: the elements holding the dof in ascending
// order, the 27 stencil entries of each; an entry between two shared dofs is emitted by the lowest
// element holding both, with the contributions of all of them.  FILL = false only counts.
template <bool FILL>
__global__ void lor_csr_rows_kernel(int p, int ndofs, const int *__restrict__ gmap, const int *__restrict__ K,
                                    const int *__restrict__ ind, const double *__restrict__ V,
                                    const int64_t *__restrict__ I64, int *__restrict__ count, int *__restrict__ J,
                                    double *__restrict__ A)
{
   const int ii = blockIdx.x * blockDim.x + threadIdx.x;
   if (ii >= ndofs) { return; }
   const int n1 = p + 1, nv = n1 * n1 * n1;
   const int i_off = K[ii], i_ne = K[ii + 1] - i_off;
   int n = 0;
   const int64_t base = FILL ? I64[ii] : 0;
   for (int k = 0; k < i_ne; k++)
   {
      const int iE = lor_decode(ind[i_off + k]);
      const int e = iE / nv, i_el = iE - e * nv;
      const int ix = i_el % n1, iy = (i_el / n1) % n1, iz = i_el / (n1 * n1);
      for (int j = 0; j < 27; j++)
      {
         const int jx = ix + j % 3 - 1, jy = iy + (j / 3) % 3 - 1, jz = iz + j / 9 - 1;
         if (jx < 0 || jx > p || jy < 0 || jy > p || jz < 0 || jz > p) { continue; }
         const int j_el = jx + n1 * (jy + n1 * jz);
         const int jj = lor_decode(gmap[(int64_t)e * nv + j_el]);
         const int j_off = K[jj], j_ne = K[jj + 1] - j_off;
         if (i_ne == 1 || j_ne == 1)
         {
            if (FILL) { J[base + n] = jj; A[base + n] = V[((int64_t)e * nv + i_el) * 27 + j]; }
            n++;
            continue;
         }
         // both lists ascend with the element number: the first common element is the minimum
         int min_e = -1;
         for (int a = 0; a < i_ne && min_e < 0; a++)
         {
            const int ea = lor_decode(ind[i_off + a]) / nv;
            for (int b = 0; b < j_ne; b++)
            {
               if (lor_decode(ind[j_off + b]) / nv == ea) { min_e = ea; break; }
            }
         }
         if (e != min_e) { continue; }
         if (FILL)
         {
            double val = 0.0;
            for (int a = 0; a < i_ne; a++)
            {
               const int iE2 = lor_decode(ind[i_off + a]);
               const int e2 = iE2 / nv, i2 = iE2 - e2 * nv;
               for (int b = 0; b < j_ne; b++)
               {
                  const int jE2 = lor_decode(ind[j_off + b]);
                  if (jE2 / nv != e2) { continue; }
                  const int j2 = jE2 - e2 * nv;
                  const int dx = j2 % n1 - i2 % n1, dy = (j2 / n1) % n1 - (i2 / n1) % n1,
                            dz = j2 / (n1 * n1) - i2 / (n1 * n1);
                  val += V[((int64_t)e2 * nv + i2) * 27 + (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)];
               }
            }
            J[base + n] = jj;
            A[base + n] = val;
         }
         n++;
      }
   }
   if (!FILL) { count[ii] = n; }
}

__global__ void lor_offsets_to_int_kernel(int n, const int64_t *__restrict__ in, int *__restrict__ out)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { out[i] = (int)in[i]; }
}

__global__ void lor_mark_kernel(int n, const int *__restrict__ ess, unsigned char *__restrict__ mark)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { mark[ess[i]] = 1; }
}

// EliminateBC(ess, DIAG_KEEP): rows and columns of the essential dofs zeroed, the diagonal kept
__global__ void lor_eliminate_bc_kernel(int ndofs, const int *__restrict__ I, const int *__restrict__ J,
                                        const unsigned char *__restrict__ mark, double *__restrict__ A)
{
   const int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= ndofs) { return; }
   const bool mr = mark[r];
   for (int k = I[r]; k < I[r + 1]; k++)
   {
      const int c = J[k];
      if (c != r && (mr || mark[c])) { A[k] = 0.0; }
   }
}

// y = A x, four lanes per row
__global__ void lor_spmv_kernel(int ndofs, const int *__restrict__ I, const int *__restrict__ J,
                                const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ y)
{
   const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int r = (int)(t >> 2), l = (int)(t & 3);
   double s = 0.0;
   if (r < ndofs)
   {
      for (int k = I[r] + l; k < I[r + 1]; k += 4) { s += A[k] * x[J[k]]; }
   }
   s += __shfl_xor_sync(0xffffffffu, s, 1);
   s += __shfl_xor_sync(0xffffffffu, s, 2);
   if (r < ndofs && l == 0) { y[r] = s; }
}

static void lor_free(b2f_lor *L)
{
   if (!L) { return; }
   cudaFree(L->I); cudaFree(L->J); cudaFree(L->A); cudaFree(L->V);
   delete L;
}

} // namespace b2f

using namespace b2f;

extern "C" {

int b2f_lor_vertex_coordinates(b2f_ctx *ctx, int dim, int p, int node_nd1d, int64_t ne, const double *Bn_h,
                               const double *nodes_e, double *x_vert)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx && Bn_h && (ne == 0 || (nodes_e && x_vert)), "b2f_lor_vertex_coordinates: NULL argument");
   B2F_REQUIRE(dim == 3, "b2f_lor_vertex_coordinates: dim = %d (3D only)", dim);
   B2F_REQUIRE(p >= 1 && p <= 16 && node_nd1d >= 2 && node_nd1d <= 17, "b2f_lor_vertex_coordinates: p = %d, nd = %d",
               p, node_nd1d);
   if (ne == 0) { return B2F_OK; }
   const int n1 = p + 1;
   double *Bn = nullptr;
   B2F_CUDA(cudaMalloc(&Bn, sizeof(double) * n1 * node_nd1d));
   int rc = B2F_OK;
   if (cudaMemcpyAsync(Bn, Bn_h, sizeof(double) * n1 * node_nd1d, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
   {
      set_error("b2f_lor_vertex_coordinates: upload failed");
      rc = B2F_ERR_CUDA;
   }
   if (rc == B2F_OK)
   {
      const int64_t n = ne * n1 * n1 * n1;
      lor_vertex_coordinates_kernel<<<(unsigned)cdiv(n, 128), 128, 0, ctx->stream>>>(n1, node_nd1d, ne, Bn, nodes_e,
                                                                                     x_vert);
      ctx->launches++;
      if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)
      {
         set_error("b2f_lor_vertex_coordinates: kernel failed");
         rc = B2F_ERR_CUDA;
      }
   }
   cudaFree(Bn);
   return rc;
}

static int lor_assemble_impl(b2f_ctx *ctx, const b2f_restriction *r, int p, const double *x_vert, const double *c_mass,
                             int64_t n_mass, const double *c_diff, int64_t n_diff, const int *ess, int n_ess,
                             int keep_sparse_ij, b2f_lor *L)
{
   const int n1 = p + 1, nv = n1 * n1 * n1;
   const int64_t ne = r->ne;
   L->ctx = ctx; L->p = p; L->ne = ne; L->ndofs = r->ndofs;
   const int nd = r->ndofs;
   B2F_CUDA(cudaMalloc(&L->I, sizeof(int) * ((size_t)nd + 1)));
   if (ne == 0 || nd == 0)
   {
      B2F_CUDA(cudaMemsetAsync(L->I, 0, sizeof(int) * ((size_t)nd + 1), ctx->stream));
      B2F_CUDA(cudaStreamSynchronize(ctx->stream));
      return B2F_OK;
   }
   B2F_CUDA(cudaMalloc(&L->V, sizeof(double) * (size_t)ne * nv * 27));
   // macro elements per CTA: about 128 sub-elements, at most 96 KB of stencil array
   const size_t per_el = sizeof(double) * nv * 27;
   int neb = 128 / (p * p * p);
   if (neb < 1) { neb = 1; }
   while (neb > 1 && neb * per_el > 96 * 1024) { neb--; }
   const size_t smem = neb * per_el;
   B2F_REQUIRE((int)smem <= ctx->max_smem_optin, "b2f_lor_assemble_h1: p = %d needs %zu bytes of shared memory", p, smem);
   B2F_CUDA(cudaFuncSetAttribute(lor_h1_assemble_3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   lor_h1_assemble_3d_kernel<<<(unsigned)cdiv(ne, neb), 128, smem, ctx->stream>>>(p, neb, ne, x_vert, c_mass,
                                                                                  n_mass == 1, c_diff, n_diff == 1,
                                                                                  L->V);
   B2F_LAUNCH_CHECK(ctx);
   // rows: count, scan, fill
   int *count = nullptr;
   int64_t *I64 = nullptr;
   void *tmp = nullptr;
   unsigned char *mark = nullptr;
   int rc = B2F_OK;
   auto cleanup = [&]() { cudaFree(count); cudaFree(I64); cudaFree(tmp); cudaFree(mark); };
#define LOR_CUDA(call)                                                                              \
   if (rc == B2F_OK && (call) != cudaSuccess)                                                        \
   {                                                                                                 \
      set_error("b2f_lor_assemble_h1: %s failed: %s", #call, cudaGetErrorString(cudaGetLastError())); \
      rc = B2F_ERR_CUDA;                                                                             \
   }
   LOR_CUDA(cudaMalloc(&count, sizeof(int) * ((size_t)nd + 1)));
   LOR_CUDA(cudaMalloc(&I64, sizeof(int64_t) * ((size_t)nd + 1)));
   LOR_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * ((size_t)nd + 1), ctx->stream));
   if (rc == B2F_OK)
   {
      lor_csr_rows_kernel<false><<<(unsigned)cdiv(nd, 128), 128, 0, ctx->stream>>>(
         p, nd, r->gather_map, r->offsets, r->indices, L->V, nullptr, count, nullptr, nullptr);
      ctx->launches++;
      LOR_CUDA(cudaGetLastError());
   }
   size_t tmp_bytes = 0;
   LOR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, count, I64, nd + 1, ctx->stream));
   LOR_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
   LOR_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, count, I64, nd + 1, ctx->stream));
   int64_t nnz = 0;
   LOR_CUDA(cudaMemcpyAsync(&nnz, I64 + nd, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
   LOR_CUDA(cudaStreamSynchronize(ctx->stream));
   if (rc == B2F_OK && nnz > 2147483647LL)
   {
      set_error("b2f_lor_assemble_h1: %lld nonzeros do not fit the 32-bit row offsets of a SparseMatrix",
                (long long)nnz);
      rc = B2F_ERR_ARG;
   }
   L->nnz = nnz;
   LOR_CUDA(cudaMalloc(&L->J, sizeof(int) * (size_t)(nnz ? nnz : 1)));
   LOR_CUDA(cudaMalloc(&L->A, sizeof(double) * (size_t)(nnz ? nnz : 1)));
   if (rc == B2F_OK)
   {
      lor_offsets_to_int_kernel<<<(unsigned)cdiv(nd + 1, 256), 256, 0, ctx->stream>>>(nd + 1, I64, L->I);
      lor_csr_rows_kernel<true><<<(unsigned)cdiv(nd, 128), 128, 0, ctx->stream>>>(
         p, nd, r->gather_map, r->offsets, r->indices, L->V, I64, nullptr, L->J, L->A);
      ctx->launches += 2;
      LOR_CUDA(cudaGetLastError());
   }
   if (rc == B2F_OK && n_ess > 0)
   {
      LOR_CUDA(cudaMalloc(&mark, (size_t)nd));
      LOR_CUDA(cudaMemsetAsync(mark, 0, (size_t)nd, ctx->stream));
      if (rc == B2F_OK)
      {
         lor_mark_kernel<<<(unsigned)cdiv(n_ess, 256), 256, 0, ctx->stream>>>(n_ess, ess, mark);
         lor_eliminate_bc_kernel<<<(unsigned)cdiv(nd, 128), 128, 0, ctx->stream>>>(nd, L->I, L->J, mark, L->A);
         ctx->launches += 2;
         LOR_CUDA(cudaGetLastError());
      }
   }
   LOR_CUDA(cudaStreamSynchronize(ctx->stream));
#undef LOR_CUDA
   cleanup();
   if (rc == B2F_OK && !keep_sparse_ij) { cudaFree(L->V); L->V = nullptr; }
   return rc;
}

int b2f_lor_assemble_h1(b2f_ctx *ctx, const b2f_restriction *r, int dim, int p, const double *x_vert,
                        const double *c_mass, int64_t n_mass, const double *c_diff, int64_t n_diff, const int *ess,
                        int n_ess, int keep_sparse_ij, b2f_lor **out)
{
   B2F_ON(ctx);
   B2F_REQUIRE(ctx && r && out && c_mass && c_diff, "b2f_lor_assemble_h1: NULL argument");
   B2F_REQUIRE(dim == 3, "b2f_lor_assemble_h1: dim = %d (3D only)", dim);
   B2F_REQUIRE(p >= 1 && p <= 8, "b2f_lor_assemble_h1: order %d (the reference has kernels for 1..8)", p);
   B2F_REQUIRE(r->vdim == 1, "b2f_lor_assemble_h1: vdim = %d", r->vdim);
   const int64_t nv = (int64_t)(p + 1) * (p + 1) * (p + 1);
   B2F_REQUIRE(r->dof == nv, "b2f_lor_assemble_h1: the restriction has %d dofs per element, order %d has %lld", r->dof,
               p, (long long)nv);
   B2F_REQUIRE(r->ne == 0 || x_vert, "b2f_lor_assemble_h1: NULL vertex coordinates");
   B2F_REQUIRE((n_mass == 1 || n_mass == nv * r->ne) && (n_diff == 1 || n_diff == nv * r->ne),
               "b2f_lor_assemble_h1: coefficient sizes %lld, %lld (1 or %lld)", (long long)n_mass, (long long)n_diff,
               (long long)(nv * r->ne));
   B2F_REQUIRE(n_ess == 0 || ess, "b2f_lor_assemble_h1: NULL essential dof list");
   B2F_REQUIRE((int64_t)r->ne * nv < 2147483647LL, "b2f_lor_assemble_h1: E-vector too long for 32-bit indices");
   b2f_lor *L = new (std::nothrow) b2f_lor;
   B2F_REQUIRE(L, "b2f_lor_assemble_h1: out of memory");
   const int rc = lor_assemble_impl(ctx, r, p, x_vert, c_mass, n_mass, c_diff, n_diff, ess, n_ess, keep_sparse_ij, L);
   if (rc != B2F_OK)
   {
      lor_free(L);
      return rc;
   }
   *out = L;
   return B2F_OK;
}

int b2f_lor_csr(const b2f_lor *L, int *n_rows, int64_t *nnz, const int **I, const int **J, const double **A)
{
   B2F_REQUIRE(L, "b2f_lor_csr: NULL argument");
   if (n_rows) { *n_rows = L->ndofs; }
   if (nnz) { *nnz = L->nnz; }
   if (I) { *I = L->I; }
   if (J) { *J = L->J; }
   if (A) { *A = L->A; }
   return B2F_OK;
}

int b2f_lor_sparse_ij(const b2f_lor *L, const double **V, int64_t *size)
{
   B2F_REQUIRE(L && V, "b2f_lor_sparse_ij: NULL argument");
   B2F_REQUIRE(L->V || L->ne == 0, "b2f_lor_sparse_ij: the stencil array was not kept (keep_sparse_ij = 0)");
   *V = L->V;
   if (size) { *size = L->ne * 27 * (int64_t)(L->p + 1) * (L->p + 1) * (L->p + 1); }
   return B2F_OK;
}

int b2f_lor_mult(b2f_lor *L, const double *x, double *y)
{
   B2F_ON(L ? L->ctx : nullptr);
   B2F_REQUIRE(L && (L->ndofs == 0 || (x && y)), "b2f_lor_mult: NULL argument");
   if (L->ndofs == 0) { return B2F_OK; }
   b2f_ctx *ctx = L->ctx;
   lor_spmv_kernel<<<(unsigned)cdiv(4 * (int64_t)L->ndofs, 128), 128, 0, ctx->stream>>>(L->ndofs, L->I, L->J, L->A, x, y);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

void b2f_lor_destroy(b2f_lor *L)
{
   B2F_ON(L ? L->ctx : nullptr);
   lor_free(L);
}

} // extern "C"
