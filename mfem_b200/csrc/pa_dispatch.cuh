// mfem_b200/csrc/pa_dispatch.cuh -- (D1D,Q1D) -> kernel instantiation tables.
// The role of MFEM's KernelDispatchTable (fem/kernel_dispatch.hpp:94-184), without the
// thread-per-element fallback: every CEED BP order p = 1..8 has a specialised kernel
// (MFEM's tables miss (9,10), i.e. p = 8, for both mass and diffusion).
#pragma once
#include "pa_apply.cuh"

namespace b2f
{

// elements per CTA: fill ~half an SM's shared memory (two CTAs per SM), <= 288 threads
constexpr int neb_3d(int kind, int D, int Q)
{
   const int nqd = kind == B2F_KIND_DIFFUSION ? 6 : 1;
   const int na = kind == B2F_KIND_DIFFUSION ? 2 : 1;
   const int nc = kind == B2F_KIND_DIFFUSION ? 3 : 1;
   const int per = 8 * (nqd * Q * Q * Q + na * D * D * Q + nc * D * Q * Q);
   int n = (110 * 1024) / per;
   const int cap = 288 / (Q * Q);
   if (n > cap) { n = cap; }
   if (n > 32) { n = 32; }
   if (n < 1) { n = 1; }
   return n;
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
      return launch_3d<KIND, D_, Q_, neb_3d(KIND, D_, Q_)>(pa, x, y, idx);
#define B2F_CASE_2D(KIND, D_, Q_)                                                      \
   case (D_) * 100 + (Q_):                                                             \
      return launch_2d<KIND, D_, Q_, neb_2d(KIND, D_, Q_)>(pa, x, y, idx);

inline int unsupported(const b2f_pa *pa)
{
   set_error("no PA kernel for kind=%d dim=%d D1D=%d Q1D=%d (supported: Q1D in {D1D, D1D+1, "
             "D1D+2}, D1D = 2..9)", pa->kind, pa->dim, pa->d1d, pa->q1d);
   return B2F_ERR_UNSUPPORTED;
}

} // namespace b2f
