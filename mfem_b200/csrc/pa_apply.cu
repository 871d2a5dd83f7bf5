// mfem_b200/csrc/pa_apply.cu -- entry point of the PA operator kernels.
#include "pa_dispatch.cuh"
namespace b2f
{
int pa_apply_3d_diffusion_2(b2f_pa *, const double *, double *, const int *, double *);
int pa_apply_3d_diffusion_6(b2f_pa *, const double *, double *, const int *, double *);
int pa_apply_3d_diffusion_8(b2f_pa *, const double *, double *, const int *, double *);
int pa_apply_3d_mass_2(b2f_pa *, const double *, double *, const int *, double *);
int pa_apply_3d_mass_6(b2f_pa *, const double *, double *, const int *, double *);
int pa_apply_3d_mass_8(b2f_pa *, const double *, double *, const int *, double *);

int pa_apply_generic(b2f_pa *pa, const double *x, double *y, const int *idx); // pa_apply_generic.cu

// (D1D, Q1D) covered by the specialised kernels
static bool in_tables(const b2f_pa *pa)
{
   const int D = pa->d1d, Q = pa->q1d;
   return D >= 2 && D <= 9 && Q >= D && Q <= D + 2;
}

int pa_apply(b2f_pa *pa, const double *x, double *y, const int *idx, double *dot_acc)
{
   if (pa->ne == 0) { return B2F_OK; }
   // full-matrix coefficients, sizes outside the tables, and vector spaces where the 3D even-odd
   // kernel does not apply (2D, a basis without the symmetry): run-time-sized kernel
   const bool vec_generic = pa->vd > 1 && !pa_apply_has_fused_dot(pa);
   if (!pa->symmetric || !in_tables(pa) || vec_generic)
   {
      if (dot_acc || pa->rb != 0 || pa->re >= 0)
      {
         set_error("pa_apply: this form needs the even-odd kernel (symmetric coefficient, D1D = 2..9, "
                   "Q1D = D1D..D1D+2)");
         return B2F_ERR_UNSUPPORTED;
      }
      return pa_apply_generic(pa, x, y, idx);
   }
   if (dot_acc && !pa_apply_has_fused_dot(pa))
   {
      set_error("pa_apply: fused dot not available for this integrator");
      return B2F_ERR_UNSUPPORTED;
   }
   if (pa->dim == 2) { return pa_apply_2d(pa, x, y, idx); }
   const int d = pa->d1d;
   if (pa->kind == B2F_KIND_DIFFUSION)
   {
      return d <= 5 ? pa_apply_3d_diffusion_2(pa, x, y, idx, dot_acc)
             : d <= 7 ? pa_apply_3d_diffusion_6(pa, x, y, idx, dot_acc)
             : pa_apply_3d_diffusion_8(pa, x, y, idx, dot_acc);
   }
   return d <= 5 ? pa_apply_3d_mass_2(pa, x, y, idx, dot_acc)
          : d <= 7 ? pa_apply_3d_mass_6(pa, x, y, idx, dot_acc)
          : pa_apply_3d_mass_8(pa, x, y, idx, dot_acc);
}
bool pa_has_fused_dot(const b2f_pa *pa) { return pa_apply_has_fused_dot(pa); }
} // namespace b2f
