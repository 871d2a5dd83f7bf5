// 2D diffusion and mass PA kernels
#include "pa_dispatch.cuh"
namespace b2f
{
static int diff2d(b2f_pa *pa, const double *x, double *y, const int *idx)
{
   switch (pa->d1d * 100 + pa->q1d)
   {
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 2, 2)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 2, 3)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 2, 4)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 3, 3)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 3, 4)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 3, 5)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 4, 4)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 4, 5)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 4, 6)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 5, 5)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 5, 6)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 5, 7)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 6, 6)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 6, 7)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 6, 8)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 7, 7)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 7, 8)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 7, 9)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 8, 8)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 8, 9)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 8, 10)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 9, 9)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 9, 10)
      B2F_CASE_2D(B2F_KIND_DIFFUSION, 9, 11)
   }
   return unsupported(pa);
}
static int mass2d(b2f_pa *pa, const double *x, double *y, const int *idx)
{
   switch (pa->d1d * 100 + pa->q1d)
   {
      B2F_CASE_2D(B2F_KIND_MASS, 2, 2)
      B2F_CASE_2D(B2F_KIND_MASS, 2, 3)
      B2F_CASE_2D(B2F_KIND_MASS, 2, 4)
      B2F_CASE_2D(B2F_KIND_MASS, 3, 3)
      B2F_CASE_2D(B2F_KIND_MASS, 3, 4)
      B2F_CASE_2D(B2F_KIND_MASS, 3, 5)
      B2F_CASE_2D(B2F_KIND_MASS, 4, 4)
      B2F_CASE_2D(B2F_KIND_MASS, 4, 5)
      B2F_CASE_2D(B2F_KIND_MASS, 4, 6)
      B2F_CASE_2D(B2F_KIND_MASS, 5, 5)
      B2F_CASE_2D(B2F_KIND_MASS, 5, 6)
      B2F_CASE_2D(B2F_KIND_MASS, 5, 7)
      B2F_CASE_2D(B2F_KIND_MASS, 6, 6)
      B2F_CASE_2D(B2F_KIND_MASS, 6, 7)
      B2F_CASE_2D(B2F_KIND_MASS, 6, 8)
      B2F_CASE_2D(B2F_KIND_MASS, 7, 7)
      B2F_CASE_2D(B2F_KIND_MASS, 7, 8)
      B2F_CASE_2D(B2F_KIND_MASS, 7, 9)
      B2F_CASE_2D(B2F_KIND_MASS, 8, 8)
      B2F_CASE_2D(B2F_KIND_MASS, 8, 9)
      B2F_CASE_2D(B2F_KIND_MASS, 8, 10)
      B2F_CASE_2D(B2F_KIND_MASS, 9, 9)
      B2F_CASE_2D(B2F_KIND_MASS, 9, 10)
      B2F_CASE_2D(B2F_KIND_MASS, 9, 11)
   }
   return unsupported(pa);
}
int pa_apply_2d(b2f_pa *pa, const double *x, double *y, const int *idx)
{
   return pa->kind == B2F_KIND_DIFFUSION ? diff2d(pa, x, y, idx) : mass2d(pa, x, y, idx);
}
} // namespace b2f
