// 3D diffusion PA kernels, D1D = 6..7 (split for parallel compilation)
#include "pa_dispatch.cuh"
namespace b2f
{
int pa_apply_3d_diffusion_6(b2f_pa *pa, const double *x, double *y, const int *idx, double *dot_acc)
{
   switch (pa->d1d * 100 + pa->q1d)
   {
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 6, 6)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 6, 7)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 6, 8)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 7, 7)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 7, 8)
      B2F_CASE_3D(B2F_KIND_DIFFUSION, 7, 9)
   }
   return unsupported(pa);
}
} // namespace b2f
