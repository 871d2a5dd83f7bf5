// mfem_b200/csrc/pa_apply.cu -- entry point of the PA operator kernels.
#include "pa_dispatch.cuh"
namespace b2f
{
int pa_apply_3d_diffusion_2(b2f_pa *, const double *, double *, const int *);
int pa_apply_3d_diffusion_6(b2f_pa *, const double *, double *, const int *);
int pa_apply_3d_diffusion_8(b2f_pa *, const double *, double *, const int *);
int pa_apply_3d_mass_2(b2f_pa *, const double *, double *, const int *);
int pa_apply_3d_mass_6(b2f_pa *, const double *, double *, const int *);
int pa_apply_3d_mass_8(b2f_pa *, const double *, double *, const int *);

int pa_apply(b2f_pa *pa, const double *x, double *y, const int *idx)
{
   if (pa->ne == 0) { return B2F_OK; }
   if (!pa->symmetric)
   {
      set_error("non-symmetric matrix coefficients have no apply kernel yet");
      return B2F_ERR_UNSUPPORTED;
   }
   if (pa->dim == 2) { return pa_apply_2d(pa, x, y, idx); }
   const int d = pa->d1d;
   if (pa->kind == B2F_KIND_DIFFUSION)
   {
      return d <= 5 ? pa_apply_3d_diffusion_2(pa, x, y, idx)
             : d <= 7 ? pa_apply_3d_diffusion_6(pa, x, y, idx)
             : pa_apply_3d_diffusion_8(pa, x, y, idx);
   }
   return d <= 5 ? pa_apply_3d_mass_2(pa, x, y, idx)
          : d <= 7 ? pa_apply_3d_mass_6(pa, x, y, idx)
          : pa_apply_3d_mass_8(pa, x, y, idx);
}
} // namespace b2f
