// 3D mass PA kernels, D1D = 2..5 (split for parallel compilation)
#include <cstdlib>
#include "pa_dispatch.cuh"
#include "pa_apply_eo_lean.cuh"
namespace b2f
{
int pa_apply_3d_mass_2(b2f_pa *pa, const double *x, double *y, const int *idx, double *dot_acc)
{
   // experiment hook (tools/opbench.py --variants v60): short-prologue form of the default kernel
   // for (D1D,Q1D) = (5,6), pa_apply_eo_lean.cuh; not yet validated on hardware
   if (pa->d1d == 5 && pa->q1d == 6 && idx && !dot_acc && pa->vd == 1)
   {
      const char *s = getenv("B2F_VARIANT");
      if (s && atoi(s) == 60)
      {
         constexpr int K = B2F_KIND_MASS, N = neb_eo(K, 5, 6, 1);
         return launch_3d_eo_lean<K, 5, 6, N, minb_eo(K, 5, 6, N)>(pa, x, y, idx);
      }
   }
   switch (pa->d1d * 100 + pa->q1d)
   {
      B2F_CASE_3D(B2F_KIND_MASS, 2, 2)
      B2F_CASE_3D(B2F_KIND_MASS, 2, 3)
      B2F_CASE_3D(B2F_KIND_MASS, 2, 4)
      B2F_CASE_3D(B2F_KIND_MASS, 3, 3)
      B2F_CASE_3D(B2F_KIND_MASS, 3, 4)
      B2F_CASE_3D(B2F_KIND_MASS, 3, 5)
      B2F_CASE_3D(B2F_KIND_MASS, 4, 4)
      B2F_CASE_3D(B2F_KIND_MASS, 4, 5)
      B2F_CASE_3D(B2F_KIND_MASS, 4, 6)
      B2F_CASE_3D(B2F_KIND_MASS, 5, 5)
      B2F_CASE_3D(B2F_KIND_MASS, 5, 6)
      B2F_CASE_3D(B2F_KIND_MASS, 5, 7)
   }
   return unsupported(pa);
}
} // namespace b2f
