// 3D mass PA kernels, D1D = 8..9 (split for parallel compilation)
#include "pa_dispatch.cuh"
namespace b2f
{
int pa_apply_3d_mass_8(b2f_pa *pa, const double *x, double *y, const int *idx, double *dot_acc)
{
   switch (pa->d1d * 100 + pa->q1d)
   {
      B2F_CASE_3D(B2F_KIND_MASS, 8, 8)
      B2F_CASE_3D(B2F_KIND_MASS, 8, 9)
      B2F_CASE_3D(B2F_KIND_MASS, 8, 10)
      B2F_CASE_3D(B2F_KIND_MASS, 9, 9)
      B2F_CASE_3D(B2F_KIND_MASS, 9, 10)
      B2F_CASE_3D(B2F_KIND_MASS, 9, 11)
   }
   return unsupported(pa);
}
} // namespace b2f
