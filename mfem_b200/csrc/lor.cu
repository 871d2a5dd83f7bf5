// mfem_b200/csrc/lor.cu -- low-order-refined (LOR) assembly of H1 diffusion + mass forms in 2D and 3D.
//
// What it replaces: BatchedLORAssembly::Assemble for H1 spaces (fem/lor/lor_batched.cpp:712-734):
//   FormLORVertexCoordinates (lor_batched.cpp:75-112), BatchedLOR_H1::Assemble2D / Assemble3D
//   (fem/lor/lor_h1_impl.hpp:19-182 / 184-455), FillI / FillJAndData (lor_batched.cpp:143-369) and
//   SparseMatrix::EliminateBC(ess, DIAG_KEEP) (linalg/sparsemat.cpp:2375-2416).  The result is the
//   CSR matrix MFEM hands to an algebraic multigrid (LORSolver<HypreBoomerAMG>, fem/lor/lor.hpp).
//
// How it differs from the reference's device path:
//   * the 8x8 matrices of a macro element's p^3 sub-hexahedra stay in shared memory and every dof
//     GATHERS its 27 stencil values from the sub-hexahedra around it in a fixed order, so there are no
//     atomics and the result is reproducible bit for bit; the reference adds with global atomics;
//   * the CSR is filled by ROWS (one warp per dof walks the elements holding it, in element order),
//     not by (element, local dof) with an atomic cursor per row: the column order inside a row is then
//     exactly the order of the reference's serial host loop, so I and J match the reference's CPU result
//     bit for bit (its own device path returns the columns of a row in a run-dependent order);
//   * the stencil index of a pair of dofs in another element is computed from their lexicographic
//     coordinates instead of searching the 27-entry map.
#include "common.cuh"

#include <cub/device/device_scan.cuh>
#include <cstdlib>
#include <new>

struct b2f_lor
{
   b2f_ctx *ctx = nullptr;
   int dim = 3, p = 0, ndofs = 0;
   int64_t ne = 0, nnz = 0;
   int *I = nullptr, *J = nullptr;
   double *A = nullptr;
   double *V = nullptr; // (3^dim, nd1d^dim, NE) macro-element stencil array ("sparse_ij")
   float kernel_ms = 0.f; // device time of the assembly's kernels (events on the stream; allocations excluded)
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

// 2D: X_vert(c, i, e) = sum_j Bn(ix,jx) Bn(iy,jy) nodes(j, c, e)
__global__ void lor_vertex_coordinates_2d_kernel(int n1, int nd, int64_t ne, const double *__restrict__ Bn,
                                                 const double *__restrict__ nodes, double *__restrict__ xv)
{
   const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int nv = n1 * n1, nn = nd * nd;
   if (tid >= ne * nv) { return; }
   const int64_t e = tid / nv;
   const int i = (int)(tid - e * nv);
   const int ix = i % n1, iy = i / n1;
   double s0 = 0.0, s1 = 0.0;
   const double *ne_ = nodes + e * 2 * (int64_t)nn;
   for (int jy = 0; jy < nd; jy++)
   {
      const double by = Bn[iy * nd + jy];
      for (int jx = 0; jx < nd; jx++)
      {
         const double b = by * Bn[ix * nd + jx];
         s0 += b * ne_[jx + nd * jy];
         s1 += b * ne_[nn + jx + nd * jy];
      }
   }
   xv[2 * tid] = s0;
   xv[2 * tid + 1] = s1;
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

// CTA = `neb` consecutive macro elements.  Phase 1: one thread per sub-hexahedron computes its 8x8
// matrix and leaves the lower triangle (36 values, odd pitch) in shared memory.  Phase 2: one thread per
// dof of the macro elements GATHERS its 27 stencil values from the (up to 8) sub-hexahedra around it, in a
// fixed order -- no atomics, no colouring, reproducible bit for bit -- and the warp writes the 32 finished
// rows (contiguous in V) through a private staging slab, fully coalesced.
constexpr int LOR_NT = 128, LOR_MP = 37; // threads per CTA; pitch of a sub-element's 36-entry triangle
__device__ __forceinline__ constexpr int lor_tri(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

__global__ void __launch_bounds__(LOR_NT)
lor_h1_assemble_3d_kernel(int p, int neb, int64_t ne, const double *__restrict__ xv,
                          const double *__restrict__ cm, int cm_const, const double *__restrict__ cd,
                          int cd_const, double *__restrict__ V)
{
   extern __shared__ double lsm[];
   const int n1 = p + 1, nv = n1 * n1 * n1, ns = p * p * p;
   const int64_t e0 = (int64_t)blockIdx.x * neb;
   const int nev = (int)((ne - e0) < neb ? (ne - e0) : neb);
   double *Ms = lsm;                                                       // [nev * ns][LOR_MP]
   double *stage = lsm + (size_t)neb * ns * LOR_MP + (threadIdx.x >> 5) * (32 * 27); // per warp [32][27]
   for (int s = threadIdx.x; s < nev * ns; s += LOR_NT)
   {
      const int el = s / ns;
      int r = s - el * ns;
      const int kx = r % p, ky = (r / p) % p, kz = r / (p * p);
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
      double *m = Ms + (size_t)s * LOR_MP;
#pragma unroll
      for (int ii = 0; ii < 8; ii++)
#pragma unroll
         for (int jj = 0; jj <= ii; jj++) { m[lor_tri(ii, jj)] = M[ii][jj]; }
   }
   __syncthreads();
   const int total = nev * nv;
   const int lane = threadIdx.x & 31;
   double *Vg = V + e0 * (int64_t)nv * 27;
   for (int d0 = (threadIdx.x >> 5) * 32; d0 < total; d0 += LOR_NT)
   {
      const int d = d0 + lane;
      double acc[27];
#pragma unroll
      for (int j = 0; j < 27; j++) { acc[j] = 0.0; }
      if (d < total)
      {
         const int el = d / nv, i = d - el * nv;
         const int ix = i % n1, iy = (i / n1) % n1, iz = i / (n1 * n1);
         // the sub-hexahedron that has this dof as its corner (cx, cy, cz) starts at (ix-cx, iy-cy, iz-cz)
#pragma unroll
         for (int c = 0; c < 8; c++)
         {
            const int cx = c & 1, cy = (c >> 1) & 1, cz = c >> 2;
            const int kx = ix - cx, ky = iy - cy, kz = iz - cz;
            if (kx < 0 || kx >= p || ky < 0 || ky >= p || kz < 0 || kz >= p) { continue; }
            const double *m = Ms + (size_t)(el * ns + kx + p * (ky + p * kz)) * LOR_MP;
#pragma unroll
            for (int jj = 0; jj < 8; jj++)
            {
               const int jx = jj & 1, jy = (jj >> 1) & 1, jz = jj >> 2;
               acc[(jx - cx + 1) + 3 * (jy - cy + 1) + 9 * (jz - cz + 1)] += m[lor_tri(c, jj)];
            }
         }
      }
#pragma unroll
      for (int j = 0; j < 27; j++) { stage[lane * 27 + j] = acc[j]; }
      __syncwarp();
      const int rows = (total - d0) < 32 ? (total - d0) : 32;
      double *out = Vg + (int64_t)d0 * 27;
      for (int t = lane; t < rows * 27; t += 32) { out[t] = stage[t]; }
      __syncwarp();
   }
}

// 2D twin (BatchedLOR_H1::Assemble2D, lor_h1_impl.hpp:19-182): sub-quadrilaterals, bilinear map, vertex
// quadrature (weight 1/4), 4x4 matrices (lower triangle: 10 values, pitch 11), 9-point stencil.
constexpr int LOR_MP2 = 11;

__global__ void __launch_bounds__(LOR_NT)
lor_h1_assemble_2d_kernel(int p, int neb, int64_t ne, const double *__restrict__ xv,
                          const double *__restrict__ cm, int cm_const, const double *__restrict__ cd,
                          int cd_const, double *__restrict__ V)
{
   extern __shared__ double lsm[];
   const int n1 = p + 1, nv = n1 * n1, ns = p * p;
   const int64_t e0 = (int64_t)blockIdx.x * neb;
   const int nev = (int)((ne - e0) < neb ? (ne - e0) : neb);
   double *Ms = lsm;
   double *stage = lsm + (size_t)neb * ns * LOR_MP2 + (threadIdx.x >> 5) * (32 * 9);
   for (int s = threadIdx.x; s < nev * ns; s += LOR_NT)
   {
      const int el = s / ns;
      const int r = s - el * ns;
      const int kx = r % p, ky = r / p;
      const int64_t e = e0 + el;
      double P[2][2][2], mq[2][2], dq[2][2];
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
         for (int dx = 0; dx < 2; dx++)
         {
            const int64_t v = (kx + dx) + n1 * (ky + dy) + (int64_t)nv * e;
            P[0][dx][dy] = xv[2 * v];
            P[1][dx][dy] = xv[2 * v + 1];
            mq[dx][dy] = cm_const ? cm[0] : cm[v];
            dq[dx][dy] = cd_const ? cd[0] : cd[v];
         }
      double M[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
         for (int j = 0; j < 4; j++) { M[i][j] = 0.0; }
#pragma unroll
      for (int qy = 0; qy < 2; qy++)
#pragma unroll
         for (int qx = 0; qx < 2; qx++)
         {
            const double J00 = P[0][1][qy] - P[0][0][qy], J10 = P[1][1][qy] - P[1][0][qy];
            const double J01 = P[0][qx][1] - P[0][qx][0], J11 = P[1][qx][1] - P[1][qx][0];
            const double detJ = J00 * J11 - J10 * J01;
            const double w = 0.25;
            const double wd = dq[qx][qy] * (w / detJ);
            const double E = J00 * J00 + J10 * J10, F = J00 * J01 + J10 * J11, G = J01 * J01 + J11 * J11;
            const double K[2][2] = {{wd * G, -(wd * F)}, {-(wd * F), wd * E}};
            const int q2[2] = {qx, qy};
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
               for (int b = 0; b < 2; b++)
#pragma unroll
                  for (int s2 = 0; s2 < 2; s2++)
#pragma unroll
                     for (int t2 = 0; t2 < 2; t2++)
                     {
                        int i2[2] = {q2[0], q2[1]}, j2[2] = {q2[0], q2[1]};
                        i2[a] = s2; j2[b] = t2;
                        const int ii = i2[0] + 2 * i2[1], jj = j2[0] + 2 * j2[1];
                        if (jj <= ii) { M[ii][jj] += ((s2 == t2) ? K[a][b] : -K[a][b]); }
                     }
            const int iq = qx + 2 * qy;
            M[iq][iq] += mq[qx][qy] * (w * detJ);
         }
      double *m = Ms + (size_t)s * LOR_MP2;
#pragma unroll
      for (int ii = 0; ii < 4; ii++)
#pragma unroll
         for (int jj = 0; jj <= ii; jj++) { m[lor_tri(ii, jj)] = M[ii][jj]; }
   }
   __syncthreads();
   const int total = nev * nv;
   const int lane = threadIdx.x & 31;
   double *Vg = V + e0 * (int64_t)nv * 9;
   for (int d0 = (threadIdx.x >> 5) * 32; d0 < total; d0 += LOR_NT)
   {
      const int d = d0 + lane;
      double acc[9];
#pragma unroll
      for (int j = 0; j < 9; j++) { acc[j] = 0.0; }
      if (d < total)
      {
         const int el = d / nv, i = d - el * nv;
         const int ix = i % n1, iy = i / n1;
#pragma unroll
         for (int c = 0; c < 4; c++)
         {
            const int cx = c & 1, cy = c >> 1;
            const int kx = ix - cx, ky = iy - cy;
            if (kx < 0 || kx >= p || ky < 0 || ky >= p) { continue; }
            const double *m = Ms + (size_t)(el * ns + kx + p * ky) * LOR_MP2;
#pragma unroll
            for (int jj = 0; jj < 4; jj++)
            {
               const int jx = jj & 1, jy = jj >> 1;
               acc[(jx - cx + 1) + 3 * (jy - cy + 1)] += m[lor_tri(c, jj)];
            }
         }
      }
#pragma unroll
      for (int j = 0; j < 9; j++) { stage[lane * 9 + j] = acc[j]; }
      __syncwarp();
      const int rows = (total - d0) < 32 ? (total - d0) : 32;
      double *out = Vg + (int64_t)d0 * 9;
      for (int t = lane; t < rows * 9; t += 32) { out[t] = stage[t]; }
      __syncwarp();
   }
}

// ---- CSR ------------------------------------------------------------------------------------------
__device__ __forceinline__ int lor_decode(int s) { return s >= 0 ? s : -1 - s; }

// Walk row `ii` as the reference's serial loop meets it: the elements holding the dof in ascending
// order, the 27 stencil entries of each; an entry between two shared dofs is emitted by the lowest
// element holding both, with the contributions of all of them (FillI / FillJAndData,
// lor_batched.cpp:143-369).  One WARP per row, lane = stencil entry:
//   * per incident element a, the 27 neighbour dofs nbr[a][.] and the 27 values val[a][.] of the row are
//     loaded once (coalesced) into the warp's slab of shared memory;
//   * element k emits column jj = nbr[k][lane] iff no earlier (lower) element of the row also holds jj,
//     with the sum of val[a][l] over the later elements a and slots l where nbr[a][l] == jj -- the same
//     pairs, in the same order, that the reference finds by intersecting the two dofs' element lists,
//     without those lists, without divisions, with broadcast reads of shared memory only;
//   * the emitted entries are compacted with a ballot (which keeps the (element, stencil entry) order)
//     and written as one contiguous run.
// Rows held by more than `maxe` elements (unstructured meshes can have any vertex valence) take the
// list-intersecting path.  FILL = false only counts.
constexpr int LOR_MAXE = 16;

// geometry of the stencil: S = 3^DIM entries, entry j is the offset (j%3-1, (j/3)%3-1, j/9-1)
template <int DIM>
struct LorStencil
{
   static constexpr int S = DIM == 3 ? 27 : 9;
   static constexpr int SP = DIM == 3 ? 28 : 12; // row pitch of the neighbour slab (whole int4 words)
   __device__ static __forceinline__ int nv(int n1) { return DIM == 3 ? n1 * n1 * n1 : n1 * n1; }
   // local dof reached from local dof i_el by stencil entry j, or -1 outside the element
   __device__ static __forceinline__ int neighbour(int i_el, int j, int n1)
   {
      const int p = n1 - 1;
      const int jx = i_el % n1 + j % 3 - 1, jy = (i_el / n1) % n1 + (j / 3) % 3 - 1;
      if (jx < 0 || jx > p || jy < 0 || jy > p) { return -1; }
      if (DIM == 2) { return jx + n1 * jy; }
      const int jz = i_el / (n1 * n1) + j / 9 - 1;
      if (jz < 0 || jz > p) { return -1; }
      return jx + n1 * (jy + n1 * jz);
   }
   // stencil entry that leads from local dof i2 to local dof j2 of the same element
   __device__ static __forceinline__ int entry(int i2, int j2, int n1)
   {
      const int dx = j2 % n1 - i2 % n1, dy = (j2 / n1) % n1 - (i2 / n1) % n1;
      const int dz = DIM == 3 ? j2 / (n1 * n1) - i2 / (n1 * n1) : -1;
      return (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1);
   }
};

template <int DIM, bool FILL>
__device__ __noinline__ int lor_csr_row_by_lists(int p, int ii, int lane, const int *__restrict__ gmap,
                                                  const int *__restrict__ K, const int *__restrict__ ind,
                                                  const double *__restrict__ V, int64_t base, int *__restrict__ J,
                                                  double *__restrict__ A)
{
   using St = LorStencil<DIM>;
   constexpr int S = St::S;
   const unsigned lt = (1u << lane) - 1u;
   const int n1 = p + 1, nv = St::nv(n1);
   const int i_off = K[ii], i_ne = K[ii + 1] - i_off;
   int n = 0;
   for (int k = 0; k < i_ne; k++)
   {
      const int iE = lor_decode(ind[i_off + k]);
      const int e = iE / nv, i_el = iE - e * nv;
      const int j_el = lane < S ? St::neighbour(i_el, lane, n1) : -1;
      bool emit = j_el >= 0;
      int jj = -1;
      double val = 0.0;
      if (emit)
      {
         jj = lor_decode(gmap[(int64_t)e * nv + j_el]);
         int j_off = 0, j_ne = 1;
         if (i_ne > 1) { j_off = K[jj]; j_ne = K[jj + 1] - j_off; }
         if (i_ne == 1 || j_ne == 1)
         {
            if (FILL) { val = V[((int64_t)e * nv + i_el) * S + lane]; }
         }
         else
         {
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
            emit = e == min_e;
            if (FILL && emit)
            {
               for (int a = 0; a < i_ne; a++)
               {
                  const int iE2 = lor_decode(ind[i_off + a]);
                  const int e2 = iE2 / nv, i2 = iE2 - e2 * nv;
                  for (int b = 0; b < j_ne; b++)
                  {
                     const int jE2 = lor_decode(ind[j_off + b]);
                     if (jE2 / nv != e2) { continue; }
                     val += V[((int64_t)e2 * nv + i2) * S + St::entry(i2, jE2 - e2 * nv, n1)];
                  }
               }
            }
         }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, emit);
      if (FILL && emit)
      {
         const int64_t pos = base + n + __popc(mask & lt);
         J[pos] = jj;
         A[pos] = val;
      }
      n += __popc(mask);
   }
   return n;
}

// PT = p (compile-time index arithmetic) or 0 (run-time p)
template <int DIM, int PT, bool FILL>
__global__ void __launch_bounds__(128)
lor_csr_rows_kernel(int p_rt, int ndofs, int maxe, const int *__restrict__ gmap, const int *__restrict__ K,
                    const int *__restrict__ ind, const double *__restrict__ V, const int64_t *__restrict__ I64,
                    int *__restrict__ count, int *__restrict__ J, double *__restrict__ A)
{
   using St = LorStencil<DIM>;
   constexpr int S = St::S, SP = St::SP;
   __shared__ __align__(16) int nbr_s[4][LOR_MAXE][SP];
   __shared__ double val_s[FILL ? 4 : 1][FILL ? LOR_MAXE : 1][S];
   const int ii = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
   if (ii >= ndofs) { return; }
   const int p = PT ? PT : p_rt;
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   const int i_off = K[ii], i_ne = K[ii + 1] - i_off;
   const int64_t base = FILL ? I64[ii] : 0;
   if (i_ne > maxe)
   {
      const int n = lor_csr_row_by_lists<DIM, FILL>(p, ii, lane, gmap, K, ind, V, base, J, A);
      if (!FILL && lane == 0) { count[ii] = n; }
      return;
   }
   const unsigned lt = (1u << lane) - 1u;
   const int n1 = p + 1, nv = St::nv(n1);
   int (*nbr)[SP] = nbr_s[w];
   if (i_ne == 1)
   {
      // a dof of one element only (its interior): the row is that element's stencil, as it stands
      const int iE = lor_decode(ind[i_off]);
      const int e = iE / nv, i_el = iE - e * nv;
      const int j_el = lane < S ? St::neighbour(i_el, lane, n1) : -1;
      const bool emit = j_el >= 0;
      const unsigned mask = __ballot_sync(0xffffffffu, emit);
      if (FILL && emit)
      {
         const int64_t pos = base + __popc(mask & lt);
         J[pos] = lor_decode(gmap[(int64_t)e * nv + j_el]);
         A[pos] = V[(int64_t)iE * S + lane];
      }
      if (!FILL && lane == 0) { count[ii] = __popc(mask); }
      return;
   }
   for (int a = 0; a < i_ne; a++)
   {
      const int iE = lor_decode(ind[i_off + a]);
      const int e = iE / nv, i_el = iE - e * nv;
      if (lane < SP)
      {
         const int j_el = lane < S ? St::neighbour(i_el, lane, n1) : -1;
         nbr[a][lane] = j_el >= 0 ? lor_decode(gmap[(int64_t)e * nv + j_el]) : -1;
         if (FILL && lane < S) { val_s[w][a][lane] = V[(int64_t)iE * S + lane]; }
      }
   }
   __syncwarp();
   int n = 0;
   for (int k = 0; k < i_ne; k++)
   {
      const int jj = lane < S ? nbr[k][lane] : -1;
      bool emit = jj >= 0;
      for (int a = 0; a < k; a++) // a lower element that also holds jj has emitted the entry
      {
         const int4 *row = reinterpret_cast<const int4 *>(nbr[a]);
#pragma unroll
         for (int l = 0; l < SP / 4; l++)
         {
            const int4 v = row[l];
            emit = emit && v.x != jj && v.y != jj && v.z != jj && v.w != jj;
         }
      }
      double val = 0.0;
      if (FILL)
      {
         val = val_s[w][k][lane < S ? lane : 0];
         for (int a = k + 1; a < i_ne; a++)
         {
            // the dofs of an element are distinct: at most one slot of element a holds jj
            const int4 *row = reinterpret_cast<const int4 *>(nbr[a]);
            int m = -1;
#pragma unroll
            for (int l = 0; l < SP / 4; l++)
            {
               const int4 v = row[l];
               m = v.x == jj ? 4 * l : m;
               m = v.y == jj ? 4 * l + 1 : m;
               m = v.z == jj ? 4 * l + 2 : m;
               m = v.w == jj ? 4 * l + 3 : m;
            }
            if (m >= 0 && jj >= 0) { val += val_s[w][a][m]; }
         }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, emit);
      if (FILL && emit)
      {
         const int64_t pos = base + n + __popc(mask & lt);
         J[pos] = jj;
         A[pos] = val;
      }
      n += __popc(mask);
   }
   if (!FILL && lane == 0) { count[ii] = n; }
}

template <int DIM, bool FILL>
static void lor_launch_rows(cudaStream_t st, int p, int nd, int maxe, const b2f_restriction *r, const double *V,
                            const int64_t *I64, int *count, int *J, double *A)
{
   const unsigned grid = (unsigned)cdiv(32 * (int64_t)nd, 128);
#define LOR_ROWS(P)                                                                                                  \
   case P:                                                                                                           \
      lor_csr_rows_kernel<DIM, P, FILL><<<grid, 128, 0, st>>>(p, nd, maxe, r->gather_map, r->offsets, r->indices, V, \
                                                               I64, count, J, A);                                     \
      break;
   switch (p)
   {
      LOR_ROWS(1) LOR_ROWS(2) LOR_ROWS(3) LOR_ROWS(4) LOR_ROWS(5) LOR_ROWS(6) LOR_ROWS(7) LOR_ROWS(8)
      default:
         lor_csr_rows_kernel<DIM, 0, FILL><<<grid, 128, 0, st>>>(p, nd, maxe, r->gather_map, r->offsets, r->indices, V,
                                                                 I64, count, J, A);
   }
#undef LOR_ROWS
}

__global__ void lor_offsets_to_int_kernel(int n, const int64_t *__restrict__ in, int *__restrict__ out)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { out[i] = (int)in[i]; }
}

__global__ void lor_mark_kernel(int n, int ndofs, const int *__restrict__ ess, unsigned char *__restrict__ mark)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   const int d = ess[i];
   if (d >= 0 && d < ndofs) { mark[d] = 1; } // an index outside the space is ignored, never written through
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
   B2F_REQUIRE(dim == 2 || dim == 3, "b2f_lor_vertex_coordinates: dim = %d", dim);
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
      const int64_t n = ne * n1 * n1 * (dim == 3 ? n1 : 1);
      if (dim == 3)
      {
         lor_vertex_coordinates_kernel<<<(unsigned)cdiv(n, 128), 128, 0, ctx->stream>>>(n1, node_nd1d, ne, Bn, nodes_e,
                                                                                        x_vert);
      }
      else
      {
         lor_vertex_coordinates_2d_kernel<<<(unsigned)cdiv(n, 128), 128, 0, ctx->stream>>>(n1, node_nd1d, ne, Bn,
                                                                                           nodes_e, x_vert);
      }
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

static int lor_assemble_impl(b2f_ctx *ctx, const b2f_restriction *r, int dim, int p, const double *x_vert, const double *c_mass,
                             int64_t n_mass, const double *c_diff, int64_t n_diff, const int *ess, int n_ess,
                             int keep_sparse_ij, b2f_lor *L)
{
   const int n1 = p + 1, nv = n1 * n1 * (dim == 3 ? n1 : 1), ns = p * p * (dim == 3 ? p : 1), S = dim == 3 ? 27 : 9;
   const int64_t ne = r->ne;
   L->ctx = ctx; L->dim = dim; L->p = p; L->ne = ne; L->ndofs = r->ndofs;
   const int nd = r->ndofs;
   B2F_CUDA(cudaMalloc(&L->I, sizeof(int) * ((size_t)nd + 1)));
   if (ne == 0 || nd == 0)
   {
      B2F_CUDA(cudaMemsetAsync(L->I, 0, sizeof(int) * ((size_t)nd + 1), ctx->stream));
      B2F_CUDA(cudaStreamSynchronize(ctx->stream));
      return B2F_OK;
   }
   B2F_CUDA(cudaMalloc(&L->V, sizeof(double) * (size_t)ne * nv * S));
   // macro elements per CTA: about one sub-element per thread
   int neb = LOR_NT / ns;
   if (neb < 1) { neb = 1; }
   const size_t smem = sizeof(double) * ((size_t)neb * ns * (dim == 3 ? LOR_MP : LOR_MP2) + (LOR_NT / 32) * 32 * S);
   B2F_REQUIRE((int)smem <= ctx->max_smem_optin, "b2f_lor_assemble_h1: p = %d needs %zu bytes of shared memory", p, smem);
   cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
   for (int i = 0; i < 4; i++) { cudaEventCreate(&ev[i]); }
   struct EvGuard { cudaEvent_t *e; ~EvGuard() { for (int i = 0; i < 4; i++) { if (e[i]) { cudaEventDestroy(e[i]); } } } } evg{ev};
   if (ev[0]) { cudaEventRecord(ev[0], ctx->stream); }
   if (dim == 3)
   {
      B2F_CUDA(cudaFuncSetAttribute(lor_h1_assemble_3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      lor_h1_assemble_3d_kernel<<<(unsigned)cdiv(ne, neb), LOR_NT, smem, ctx->stream>>>(
         p, neb, ne, x_vert, c_mass, n_mass == 1, c_diff, n_diff == 1, L->V);
   }
   else
   {
      B2F_CUDA(cudaFuncSetAttribute(lor_h1_assemble_2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      lor_h1_assemble_2d_kernel<<<(unsigned)cdiv(ne, neb), LOR_NT, smem, ctx->stream>>>(
         p, neb, ne, x_vert, c_mass, n_mass == 1, c_diff, n_diff == 1, L->V);
   }
   B2F_LAUNCH_CHECK(ctx);
   // rows: count, scan, fill.  B2F_LOR_MAXE (testing): rows held by more elements than this take the
   // list-intersecting path
   int maxe = LOR_MAXE;
   if (const char *me = getenv("B2F_LOR_MAXE")) { maxe = atoi(me) < LOR_MAXE ? atoi(me) : LOR_MAXE; }
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
      if (dim == 3) { lor_launch_rows<3, false>(ctx->stream, p, nd, maxe, r, L->V, nullptr, count, nullptr, nullptr); }
      else { lor_launch_rows<2, false>(ctx->stream, p, nd, maxe, r, L->V, nullptr, count, nullptr, nullptr); }
      ctx->launches++;
      LOR_CUDA(cudaGetLastError());
   }
   size_t tmp_bytes = 0;
   LOR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, count, I64, nd + 1, ctx->stream));
   LOR_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
   LOR_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, count, I64, nd + 1, ctx->stream));
   if (ev[1]) { cudaEventRecord(ev[1], ctx->stream); }
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
   if (ev[2]) { cudaEventRecord(ev[2], ctx->stream); }
   if (rc == B2F_OK)
   {
      lor_offsets_to_int_kernel<<<(unsigned)cdiv(nd + 1, 256), 256, 0, ctx->stream>>>(nd + 1, I64, L->I);
      if (dim == 3) { lor_launch_rows<3, true>(ctx->stream, p, nd, maxe, r, L->V, I64, nullptr, L->J, L->A); }
      else { lor_launch_rows<2, true>(ctx->stream, p, nd, maxe, r, L->V, I64, nullptr, L->J, L->A); }
      ctx->launches += 2;
      LOR_CUDA(cudaGetLastError());
   }
   if (rc == B2F_OK && n_ess > 0)
   {
      LOR_CUDA(cudaMalloc(&mark, (size_t)nd));
      LOR_CUDA(cudaMemsetAsync(mark, 0, (size_t)nd, ctx->stream));
      if (rc == B2F_OK)
      {
         lor_mark_kernel<<<(unsigned)cdiv(n_ess, 256), 256, 0, ctx->stream>>>(n_ess, nd, ess, mark);
         lor_eliminate_bc_kernel<<<(unsigned)cdiv(nd, 128), 128, 0, ctx->stream>>>(nd, L->I, L->J, mark, L->A);
         ctx->launches += 2;
         LOR_CUDA(cudaGetLastError());
      }
   }
   if (ev[3]) { cudaEventRecord(ev[3], ctx->stream); }
   LOR_CUDA(cudaStreamSynchronize(ctx->stream));
#undef LOR_CUDA
   if (rc == B2F_OK && ev[0] && ev[1] && ev[2] && ev[3])
   {
      float a = 0.f, b = 0.f;
      if (cudaEventElapsedTime(&a, ev[0], ev[1]) == cudaSuccess && cudaEventElapsedTime(&b, ev[2], ev[3]) == cudaSuccess)
      {
         L->kernel_ms = a + b;
      }
   }
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
   B2F_REQUIRE(dim == 2 || dim == 3, "b2f_lor_assemble_h1: dim = %d", dim);
   B2F_REQUIRE(p >= 1 && p <= 8, "b2f_lor_assemble_h1: order %d (the reference has kernels for 1..8)", p);
   B2F_REQUIRE(r->vdim == 1, "b2f_lor_assemble_h1: vdim = %d", r->vdim);
   const int64_t nv = (int64_t)(p + 1) * (p + 1) * (dim == 3 ? p + 1 : 1);
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
   const int rc = lor_assemble_impl(ctx, r, dim, p, x_vert, c_mass, n_mass, c_diff, n_diff, ess, n_ess, keep_sparse_ij, L);
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

int b2f_lor_kernel_ms(const b2f_lor *L, double *ms)
{
   B2F_REQUIRE(L && ms, "b2f_lor_kernel_ms: NULL argument");
   *ms = L->kernel_ms;
   return B2F_OK;
}

int b2f_lor_sparse_ij(const b2f_lor *L, const double **V, int64_t *size)
{
   B2F_REQUIRE(L && V, "b2f_lor_sparse_ij: NULL argument");
   B2F_REQUIRE(L->V || L->ne == 0, "b2f_lor_sparse_ij: the stencil array was not kept (keep_sparse_ij = 0)");
   *V = L->V;
   if (size) { *size = L->ne * (L->dim == 3 ? 27 : 9) * (int64_t)(L->p + 1) * (L->p + 1) * (L->dim == 3 ? L->p + 1 : 1); }
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
