// mfem_b200/csrc/pa_eo_cfg.cuh -- shared-memory geometry of the 3D even-odd PA kernel
// (pa_apply_eo.cuh) and the async-copy helpers it uses.
//
// The kernel replaces MFEM's SmemPADiffusionApply3D / PADiffusionApply3D
// (fem/integ/bilininteg_diffusion_kernels.hpp:796-1215), SmemPAMassApply3D / PAMassApply3D
// (fem/integ/bilininteg_mass_kernels.hpp:664-1165) and -- FUSED -- ElementRestriction::Mult /
// MultTranspose (fem/restriction.cpp:116,159) plus the E-vector zero fill
// (fem/bilinearform_ext.cpp:543-557).
//
// Geometry (measured on B200, profiles/):
//  * "Pencil" decomposition: in every contraction stage one thread owns one 1D line, loads
//    it into registers once and produces every output of every field it feeds, so each
//    shared-memory word is read once per stage.
//  * Stage order y -> z -> x with yq = qy + Q*e as the FASTEST thread and address index of
//    stages 2-4: all inter-stage buffers are read and written as consecutive addresses by
//    consecutive lanes (no bank conflicts, no transposition), the x-stage can overwrite its
//    input in place (each thread owns its column: no barrier, no staging registers), and a
//    lane's quadrature-data rows are Q contiguous doubles read with 128-bit LDS.
//  * Quadrature data reaches shared memory by TMA: one bulk copy per element into a padded
//    slab, or -- Q = 8, where a row is 64 B and eight lanes reading eight consecutive rows
//    with LDS.128 would hit two of the eight 16-byte bank groups (4-way conflict, measured
//    576 of 1448 wavefronts per p=6 element) -- tensor copies with the 64-byte swizzle
//    (chunk c of row r lands at chunk c ^ ((r >> 1) & 3)), which makes the same reads
//    conflict-free without touching the layout in HBM.
#pragma once
#include "pa_apply.cuh"
#ifndef B2F_ALIAS_MIN_D
#define B2F_ALIAS_MIN_D 8 // smallest D1D whose one-element tiles overlay the y-stage buffer on the slab tail
#endif
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
   // Q = 8: 64-byte rows, TMA tensor copies with the 64-byte swizzle (see the file header);
   // slabs are unpadded (24 576 / 4 096 B, multiples of the 512-byte swizzle period) and the
   // tile starts at a 1024-byte aligned shared-memory address found at run time.
   // Q = 4 (32-byte rows, 2-way conflict) has the analogous cure -- 32-byte swizzle (chunk ^= address
   // bit 7, i.e. qz & 1) plus 128 B of padding per element slab (an odd number of 128-byte lines per
   // slab flips the phase from one element to the next) -- implemented below (SWZ4) and parity-clean,
   // but MEASURED SLOWER (profiles/r02_opbench_q4_swizzle.txt: BP5 p=3 84.7 -> 74.3 %, BP3 p=2
   // 83.8 -> 78.9 % of the HBM roofline): a tensor copy of 32-byte rows issues sector-sized requests,
   // which costs more than the 2-way conflict it removes.  It stays off.
   static constexpr bool SWZ4 = false;
   static constexpr bool SWZ = TMA_OK && (Q == 8 || (SWZ4 && Q == 4));
   static constexpr int SWZ_PAD = (SWZ && Q == 4) ? 16 : 0; // doubles
   static constexpr int ROWS = NQD * Q2;                  // rows of Q doubles per element
   static constexpr int BOX_ROWS = ROWS <= 256 ? ROWS : ROWS / 2; // tensor-copy box (<= 256 rows)
   static constexpr int NBOX = ROWS / BOX_ROWS;
   static constexpr int DS = NQD * Q3 + ((TMA_OK && !SWZ) ? L.dpad : SWZ_PAD);
   static_assert(!SWZ || Q == 8 || ((DS * 8) % 256) == 128, "Q = 4 slabs must be an odd number of 128-byte lines");
   static constexpr int QD = NEB * DS;
   // High orders (one element per CTA, D1D >= 8) are bound by residency: 63 / 87 KB per CTA leave 3 / 2
   // CTAs per SM.  The y-stage buffer (R1 doubles) is dead between the z- and the z^T-stage -- exactly
   // while the quadrature data is consumed -- so it is overlaid on the TAIL of the element's slab:
   // the slab arrives in two bulk copies, the head at kernel entry, the last QB doubles (the region
   // shared with the buffer) as soon as the z-stage has released it, on a second mbarrier.  52 / 73 KB
   // per CTA: 4 / 3 CTAs per SM.
   static constexpr bool ALIAS = DIFF && TMA_OK && !SWZ && NEB == 1 && D >= B2F_ALIAS_MIN_D;
   static constexpr int QB = ALIAS ? ((R1 + 1) / 2) * 2 : 0;   // doubles of the slab's tail, copied late
   static_assert(!ALIAS || QB < NQD * Q3, "the aliased buffer must fit in the slab");
   static constexpr size_t SMEM = 16 + sizeof(double) * (size_t)(QD + (ALIAS ? 0 : R1) + R2) + (SWZ ? 1024 : 0);
   static_assert(T1 <= NT && T2 <= NT, "one task per thread per stage");
};

// 128-byte tensor-map descriptor as a kernel parameter (CUtensorMap, built on the host by
// make_qdata_tensor_map in pa_apply_eo.cuh); an empty struct for kernels that do not use one
template <bool ON> struct TmapArg { };
template <> struct TmapArg<true> { alignas(64) unsigned char bytes[128]; };

// one TMA tensor copy: box (Q, BOX_ROWS, 1) of the (Q, rows-per-element, NE) tensor at row r0 of
// element e, completing on the mbarrier
__device__ __forceinline__ void tensor_g2s_3d(void *dst, const void *tmap, int c0, int c1, int c2,
                                              uint64_t *bar)
{
   asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
      "r"(smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// one swizzled row of Q doubles -> registers: chunk j (16 B) of the row sits at chunk j ^ s.
// Q = 8 (64-byte swizzle): s = (row >> 1) & 3, the row counted from a 512-byte aligned address;
// Q = 4 (32-byte swizzle): s = bit 7 of the row's shared-memory address.
template <int Q>
__device__ __forceinline__ void load_row_swz(const double *__restrict__ row, int s, double (&v)[Q])
{
   const char *p = reinterpret_cast<const char *>(row);
#pragma unroll
   for (int j = 0; j < Q / 2; j++)
   {
      const double2 w = *reinterpret_cast<const double2 *>(p + ((j ^ s) << 4));
      v[2 * j] = w.x; v[2 * j + 1] = w.y;
   }
}
// swizzle phase of the row (qy, qz) of element e of the tile
template <int Q>
__device__ __forceinline__ int swz_phase(int e, int qy, int qz)
{
   return Q == 8 ? (qy >> 1) & 3 : (qz ^ e) & 1;
}

__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src)
{
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
   asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N> // wait until at most N of this thread's committed groups are still in flight
__device__ __forceinline__ void cp_async_wait_group()
{
   asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
   asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
   asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}


} // namespace b2f
