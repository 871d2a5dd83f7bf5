// 3D mass PA kernels, D1D = 2..5 (split for parallel compilation)
#include "pa_dispatch.cuh"
namespace b2f
{
int pa_apply_3d_mass_2(b2f_pa *pa, const double *x, double *y, const int *idx, double *dot_acc)
{
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
