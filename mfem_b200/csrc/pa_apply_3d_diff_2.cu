// 3D diffusion PA kernels, D1D = 2..5 (split for parallel compilation)
#include <cstdlib>
#include "pa_dispatch.cuh"
#include "pa_apply_eo_lean.cuh"
namespace b2f
{
// experiment hook: B2F_VARIANT selects a kernel variant for (D1D,Q1D) = (5,6)
static int variant()
{
   const char *s = getenv("B2F_VARIANT");
   return s ? atoi(s) : 0;
}

int pa_apply_3d_diffusion_2(b2f_pa *pa, const double *x, double *y, const int *idx, double *dot_acc)
{
   if (pa->d1d == 5 && pa->q1d == 6 && idx)
   {
      constexpr int K = B2F_KIND_DIFFUSION;
      switch (variant())
      {
         // kept selectable for A/B measurements (tools/opbench.py --variants vN)
         case 1: return launch_3d<K, 5, 6, 6, 0, 1, 1>(pa, x, y, idx);          // general kernel, TMA slab, scalar LDS
         case 4: return launch_3d<K, 5, 6, 3, 1, 3, 1>(pa, x, y, idx);          // general kernel, 128-bit rows, 3 el/CTA
         case 25: return launch_3d_pipe<K, 5, 6, 4, 3>(pa, x, y, idx);          // persistent pipelined, plain contraction
         case 40: return launch_3d_eo<K, 5, 6, 3, 4, 1, true>(pa, x, y, idx);   // persistent pipelined, even-odd
         case 42: return launch_3d_eo<K, 5, 6, 4, 3, 1, true>(pa, x, y, idx);
         // short-prologue form of the default kernel (pa_apply_eo_lean.cuh), not yet validated
         case 60: if (!dot_acc && pa->vd == 1) { return launch_3d_eo_lean<K, 5, 6, 3, 4>(pa, x, y, idx); } break;
         default: break;
      }
   }
   switch (pa->d1d * 100 + pa->q1d)
   {
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 2, 2)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 2, 3)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 2, 4)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 3, 3)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 3, 4)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 3, 5)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 4, 4)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 4, 5)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 4, 6)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 5, 5)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 5, 6)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 5, 7)
   }
   return unsupported(pa);
}
} // namespace b2f
