// mfem_b200/csrc/pa_dispatch.cuh -- (D1D,Q1D) -> kernel instantiation tables.
// The role of MFEM's KernelDispatchTable (fem/kernel_dispatch.hpp:94-184), without the
// thread-per-element fallback: every CEED BP order p = 1..8 has a specialised kernel
// (MFEM's tables miss (9,10), i.e. p = 8, for both mass and diffusion).
#pragma once
#include "pa_apply_eo.cuh"

namespace b2f
{

// shared memory of one CTA of the 3D kernel holding n elements (bytes; mirrors Cfg3D::SMEM)
constexpr int smem_3d(int kind, int D, int Q, int n)
{
   const int nqd = kind == B2F_KIND_DIFFUSION ? 6 : 1;
   const int na = kind == B2F_KIND_DIFFUSION ? 2 : 1;
   const int nc = kind == B2F_KIND_DIFFUSION ? 3 : 1;
   const int t1 = n * D * D, t2 = n * Q * D, t3 = n * Q * Q;
   const int sa = pad16(t2, D), sf = pad16(t1, D), sc = pad16(t3, 1), se = pad16(t2, 1);
   const int r1 = na * D * sa > na * Q * sf ? na * D * sa : na * Q * sf;
   const int r2 = nc * D * sc > nc * Q * se ? nc * D * sc : nc * Q * se;
   return 16 + 8 * (n * nqd * Q * Q * Q + r1 + r2);
}

// Elements per CTA.  Candidate i targets a different residency: 0 -> >= 4 CTAs per SM,
// 1 -> 3 CTAs, 2 -> 2 CTAs (of the 227 KB an SM offers), always <= 512 threads.
constexpr int neb_3d(int kind, int D, int Q, int cand)
{
   const int budget = cand == 0 ? 55 * 1024 : cand == 1 ? 74 * 1024 : 112 * 1024;
   int n = 1;
   while (n < 32 && smem_3d(kind, D, Q, n + 1) <= budget && (n + 1) * Q * Q <= 512) { n++; }
   return n;
}

// CTAs per SM the launch bounds ask for: what shared memory allows, but never fewer than
// ~80 registers per thread
constexpr int minb_3d(int kind, int D, int Q, int n)
{
   const int nt = ((n * Q * Q + 31) / 32) * 32;
   int by_smem = (227 * 1024) / (smem_3d(kind, D, Q, n) + 1024);
   const int by_regs = 65536 / (nt * 80);
   int b = by_smem < by_regs ? by_smem : by_regs;
   if (b > 8) { b = 8; }
   return b < 1 ? 1 : b;
}

// measured choice of the candidate per (kind, D, Q) (tools/opbench.py sweeps, profiles/)
constexpr int tuned_cand_3d(int kind, int D, int Q)
{
   // BP3-type (Q = D+2) diffusion, sweep of profiles/r01_sweep_*.txt
   if (kind == B2F_KIND_DIFFUSION)
   {
      if (D == 2 || D == 4) { return 1; }
      if (D == 7) { return 2; }
      return 0;
   }
   return D <= 4 ? 0 : 1;
}

constexpr int neb_2d(int kind, int D, int Q)
{
   (void)kind; (void)D;
   int n = 256 / Q;
   if (n > 64) { n = 64; }
   return n < 1 ? 1 : n;
}

#define B2F_CASE_3D(KIND, D_, Q_)                                                      \
   case (D_) * 100 + (Q_):                                                             \
      return run_3d<KIND, D_, Q_>(pa, x, y, idx, dot_acc);
#define B2F_CASE_2D(KIND, D_, Q_)                                                      \
   case (D_) * 100 + (Q_):                                                             \
      return launch_2d<KIND, D_, Q_, neb_2d(KIND, D_, Q_)>(pa, x, y, idx);

// ---- even-odd kernel (pa_apply_eo.cuh) sizing: mirrors PipeCfg::SMEM --------------------------
constexpr int smem_eo(int kind, int D, int Q, int n)
{
   const int nqd = kind == B2F_KIND_DIFFUSION ? 6 : 1;
   const int na = kind == B2F_KIND_DIFFUSION ? 2 : 1;
   const int nc = kind == B2F_KIND_DIFFUSION ? 3 : 1;
   const int yq = Q * n;
   const PipeLayout L = pipe_layout(kind, D, Q, n);
   const int sa = pad16(D * yq, L.ra), sc = pad16(Q * yq, yq);
   const int q3 = Q * Q * Q;
   const bool tma = (nqd * q3) % 2 == 0, swz = tma && Q == 8; // (PipeCfg::SWZ4 is off)
   const int dp = (tma && !swz) ? L.dpad : (swz && Q == 4) ? 16 : 0;
   const bool alias = kind == B2F_KIND_DIFFUSION && tma && !swz && n == 1 && D >= B2F_ALIAS_MIN_D; // PipeCfg::ALIAS
   return 16 + 8 * (n * (nqd * q3 + dp) + (alias ? 0 : na * D * sa) + nc * D * sc) + (swz ? 1024 : 0);
}

// Elements per CTA, candidate i: 0 -> >= 4 CTAs per SM, 1 -> 3 CTAs, 2 -> 2 CTAs; <= 512 threads
constexpr int neb_eo(int kind, int D, int Q, int cand)
{
   // mass has ~7x less quadrature data per element: smaller tiles, many more CTAs per SM
   // (its kernel is latency/barrier-bound, profiles/r01_bp1_p4_ncu.txt)
   const int budget = kind == B2F_KIND_MASS ? (cand == 0 ? 8 * 1024 : cand == 1 ? 14 * 1024 : 27 * 1024)
                      : (cand == 0 ? 55 * 1024 : cand == 1 ? 74 * 1024 : 112 * 1024);
   int n = 1;
   while (n < 32 && smem_eo(kind, D, Q, n + 1) <= budget && (n + 1) * Q * Q <= 512) { n++; }
   return n;
}

constexpr int minb_eo(int kind, int D, int Q, int n)
{
   const int nt = ((n * Q * Q + 31) / 32) * 32;
   const int by_smem = (227 * 1024) / (smem_eo(kind, D, Q, n) + 1024);
   const int by_regs = 65536 / (nt * (kind == B2F_KIND_MASS ? (D <= 7 ? 56 : 72) : 120));
   int b = by_smem < by_regs ? by_smem : by_regs;
   // mass: the kernel waits on the index -> x gather chain; resident warps hide it (measured:
   // BP1 p=4 60 -> 66 GDOF/s going from 9 to 12 CTAs per SM)
   const int cap = kind == B2F_KIND_MASS ? 16 : 12;
   if (b > cap) { b = cap; }
   return b < 1 ? 1 : b;
}

// measured choice of the candidate per (kind, D, Q) (tools/opbench.py sweeps, profiles/)
constexpr int tuned_cand_eo(int kind, int D, int Q)
{
   // profiles/r01_sweep_eo_kernel.txt: the smallest tile (>= 4 CTAs per SM) wins except at the
   // lowest orders, where a CTA of the smallest tile has too few threads
   if (kind == B2F_KIND_DIFFUSION) { return (D <= 3 && Q >= D + 1) ? 1 : 0; }
   return 1;
}

// persistent form of the fused kernel (pa_apply_3d_eo_persistent_kernel): 1 = instantiated and the
// default, -1 = not instantiated.  Measured over BP1/BP3/BP5, p = 1..8
// (profiles/r02_sweep_persistent.txt): with >= 3 resident CTAs per SM the one-shot kernel already
// overlaps one CTA's gather chain with the others' arithmetic, and the kernels are bound by the
// LSU data pipe, not by latency; the persistent form wins only where a CTA is short-lived.
constexpr int tuned_persist_eo(int kind, int D, int Q)
{
   if (kind == B2F_KIND_DIFFUSION && ((D == 3 && Q == 4) || (D == 2 && Q == 2) || (D == 8 && Q == 8))) { return 1; }
   return -1;
}

// thread-per-element mass kernel (pa_apply_tpe.cuh) by default for this D1D (Q1D = D1D + 1)?
// Measured (profiles/r02_opbench_tpe.txt, r02_opbench_tpe_p2_*.txt): BP1 p=1 43 -> 76.5 % of the HBM
// roofline, 79 % with the L2 prefetch of the next tile.  p=2: the first build (27-int index rows read
// one int per lane, Q^3 array live: 229 registers) tied with the pencil kernel at 50 %; with the index
// tile staged through shared memory (the strided row reads were 27 LSU wavefronts per element, as many
// as everything else together), the z / D / z^T stages fused per column (184 registers) and the
// prefetch it reaches 65-69 %.  Tile shapes 64x5, 96x3, 160x2, 256x1 and a variant with the index tile
// aliased into the data region (3 CTAs per SM, two more barriers and a second index fetch) were all
// slower than 128 threads x 2 CTAs.  A diffusion twin for (2,2) (BP5 p=1) was written, parity-clean,
// and dropped: 71.5 vs 72.2 %.
constexpr bool tuned_tpe_mass(int D)
{
   return D == 2 || D == 3;
}

// the even-odd kernel (and with it the fused x.(Ax) accumulation) runs for 3D integrators with
// a symmetric basis and a (D,Q) in the dispatch tables
inline bool basis_is_symmetric(const b2f_pa *pa);
inline bool pa_apply_has_fused_dot(const b2f_pa *pa)
{
   const int D = pa->d1d, Q = pa->q1d;
   return pa->dim == 3 && pa->symmetric && D >= 2 && D <= 9 && Q >= D && Q <= D + 2 &&
          basis_is_symmetric(pa) && !getenv("B2F_GENERAL_KERNEL");
}

inline bool basis_is_symmetric(const b2f_pa *pa)
{
   const int D = pa->d1d, Q = pa->q1d;
   double bmax = 0.0, gmax = 0.0, asym = 0.0;
   for (int i = 0; i < D * Q; i++) { bmax = fmax(bmax, fabs(pa->B[i])); gmax = fmax(gmax, fabs(pa->G[i])); }
   for (int q = 0; q < Q; q++)
      for (int d = 0; d < D; d++)
      {
         const int a = q + Q * d, b = (Q - 1 - q) + Q * (D - 1 - d);
         if (bmax > 0) { asym = fmax(asym, fabs(pa->B[a] - pa->B[b]) / bmax); }
         if (gmax > 0) { asym = fmax(asym, fabs(pa->G[a] + pa->G[b]) / gmax); }
      }
   return asym <= 1e-12;
}

// One (kind, D, Q).  Symmetric bases (every H1 basis on symmetric nodes with a symmetric rule:
// all CEED BPs) run the even-odd kernel: the fused L-vector form (the hot path) with three
// elements-per-CTA candidates -- the tuned table picks one, B2F_NEB_CAND = 0|1|2 overrides it
// for experiments -- and the E-vector form (AddMultPA seam) with one.  A basis without the
// symmetry falls back to the general pencil kernel of pa_apply.cuh.
template <int KIND, int D, int Q>
int run_3d(b2f_pa *pa, const double *x, double *y, const int *idx, double *dot_acc)
{
   if (!pa_apply_has_fused_dot(pa))
   {
      if (dot_acc) { set_error("pa_apply: fused dot not available for this integrator"); return B2F_ERR_UNSUPPORTED; }
      constexpr int NG = neb_3d(KIND, D, Q, 1);
      return launch_3d<KIND, D, Q, NG, 1, 1, 2>(pa, x, y, idx);
   }
   constexpr int N0 = neb_eo(KIND, D, Q, 0), N1 = neb_eo(KIND, D, Q, 1), N2 = neb_eo(KIND, D, Q, 2);
   if (!idx) { return launch_3d_eo<KIND, D, Q, N1, 1, 0>(pa, x, y, idx); }
   const char *s = getenv("B2F_NEB_CAND");
   const int cand = s ? atoi(s) : tuned_cand_eo(KIND, D, Q);
   constexpr int PS = tuned_persist_eo(KIND, D, Q);
   if (cand == 0 || (cand == 1 && N1 == N0) || (cand == 2 && N2 == N0))
   {
      return launch_3d_eo<KIND, D, Q, N0, minb_eo(KIND, D, Q, N0), 1, PS>(pa, x, y, idx, dot_acc);
   }
   if (cand == 1 || (cand == 2 && N2 == N1))
   {
      if constexpr (N1 != N0) { return launch_3d_eo<KIND, D, Q, N1, minb_eo(KIND, D, Q, N1), 1, PS>(pa, x, y, idx, dot_acc); }
   }
   if constexpr (N2 != N1 && N2 != N0) { return launch_3d_eo<KIND, D, Q, N2, minb_eo(KIND, D, Q, N2), 1, PS>(pa, x, y, idx, dot_acc); }
   return B2F_ERR_UNSUPPORTED;
}

inline int unsupported(const b2f_pa *pa)
{
   set_error("no PA kernel for kind=%d dim=%d D1D=%d Q1D=%d (supported: Q1D in {D1D, D1D+1, "
             "D1D+2}, D1D = 2..9)", pa->kind, pa->dim, pa->d1d, pa->q1d);
   return B2F_ERR_UNSUPPORTED;
}

} // namespace b2f
