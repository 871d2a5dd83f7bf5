// 3D mass PA kernels, D1D = 2..5 (split for parallel compilation)
#include "pa_dispatch.cuh"
#include "pa_apply_tpe.cuh"
#ifndef B2F_TPE_MINB
#define B2F_TPE_MINB 4 // resident CTAs the p = 1 thread-per-element kernel is compiled for
#endif
namespace b2f
{
int pa_apply_3d_mass_2(b2f_pa *pa, const double *x, double *y, const int *idx, double *dot_acc)
{
   // lowest orders, fused scalar form: one thread per element (pa_apply_tpe.cuh); B2F_TPE=0|1 overrides
   if (idx && !dot_acc && pa->vd == 1 && pa->d1d <= 3 && pa->q1d == pa->d1d + 1)
   {
      const char *s = getenv("B2F_TPE");
      const bool tpe = s ? atoi(s) != 0 : tuned_tpe_mass(pa->d1d);
      if (tpe && pa->d1d == 2) { return launch_mass_tpe<2, 3, 128, B2F_TPE_MINB>(pa, x, y, idx); }
      if (tpe && pa->d1d == 3) { return launch_mass_tpe<3, 4, 128, 2>(pa, x, y, idx); }
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
