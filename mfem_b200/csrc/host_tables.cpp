// mfem_b200/csrc/host_tables.cpp -- host-side 1D quadrature rules and basis tables.
//
// Product host code (C++), mirroring what the hot path consumes from the reference:
//   QuadratureFunctions1D::GaussLegendre / GaussLobatto   fem/intrules.cpp:620,708
//   Poly_1D::Basis (Barycentric) ctor and Eval(u,d)       fem/fe/fe_base.cpp:1811,1947
//   TensorBasisElement::GetTensorDofToQuad                fem/fe/fe_base.cpp:2697
//   IntegrationRule(irx,iry[,irz]) tensor weights         fem/intrules.cpp:43-96
// The same Newton iterations / barycentric formulas are used so that B, G, W agree with
// the reference to the last bit (checked in tests/test_host_cpu.py against fixtures
// dumped from the reference).
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <vector>

#include "../../include/b200fem.h"

namespace b2f
{
void set_error(const char *fmt, ...);

namespace
{
const double kPi = 3.14159265358979323846;

// Legendre P_n(z) and P_{n-1}(z) by the three-term recurrence
inline void legendre(int n, double z, double &pn, double &pnm1)
{
   double a = 1.0, b = z; // P_0, P_1
   for (int j = 2; j <= n; j++)
   {
      const double c = ((2 * j - 1) * z * b - (j - 1) * a) / j;
      a = b;
      b = c;
   }
   pn = b;
   pnm1 = a;
}
} // namespace

void gauss_legendre(int np, double *x, double *w)
{
   switch (np)
   {
      case 1: x[0] = 0.5; w[0] = 1.0; return;
      case 2:
         x[0] = 0.21132486540518711775; x[1] = 0.78867513459481288225;
         w[0] = w[1] = 0.5;
         return;
      case 3:
         x[0] = 0.11270166537925831148; x[1] = 0.5; x[2] = 0.88729833462074168852;
         w[0] = w[2] = 5. / 18.; w[1] = 4. / 9.;
         return;
   }
   const int n = np;
   for (int i = 1; i <= (n + 1) / 2; i++)
   {
      // Newton on P_n in [-1,1], started from the Chebyshev-like guess; the last step is
      // folded into the map to [0,1] to avoid cancellation.
      double z = std::cos(kPi * (i - 0.25) / (n + 0.5));
      double xi = 0.0, dp = 0.0;
      for (bool last = false;;)
      {
         double pn, pm;
         legendre(n, z, pn, pm);
         dp = n * (z * pn - pm) / (z * z - 1);
         if (last) { break; }
         const double dz = pn / dp;
         if (std::fabs(dz) < 1e-16)
         {
            last = true;
            xi = ((1 - z) + dz) / 2;
         }
         z -= dz;
      }
      x[i - 1] = xi;
      x[n - i] = 1 - xi;
      w[i - 1] = w[n - i] = 1. / (4 * xi * (1 - xi) * dp * dp);
   }
}

void gauss_lobatto(int np, double *x, double *w)
{
   if (np == 1) { x[0] = 0.5; w[0] = 1.0; return; }
   x[0] = 0.0; x[np - 1] = 1.0;
   w[0] = w[np - 1] = 1.0 / (np * (np - 1));
   for (int i = 1; i <= (np - 1) / 2; ++i)
   {
      // interior points: zeros of P'_{np-1}; Newton on (x^2-1) P'_{np-1}
      double xi = std::sin(kPi * ((double)(i) / (np - 1) - 0.5));
      double zi = 0.0, pl = 0.0;
      for (bool last = false;;)
      {
         double pm;
         if (np - 1 >= 2) { legendre(np - 1, xi, pl, pm); }
         else { pl = xi; pm = 1.0; }
         if (last) { break; }
         const double dx = (xi * pl - pm) / (np * pl);
         if (std::fabs(dx) < 1e-16)
         {
            last = true;
            zi = ((1.0 + xi) - dx) / 2;
         }
         xi -= dx;
      }
      x[i] = zi;
      w[i] = (double)(1.0 / (np * (np - 1) * pl * pl));
      x[np - 1 - i] = 1.0 - zi;
      w[np - 1 - i] = w[i];
   }
}

// Lagrange basis on `nodes` (increasing) and its derivative at y, barycentric form
// organised around the node nearest to y so that no term is ever divided by ~0.
static void lagrange_eval(const std::vector<double> &x, const std::vector<double> &bw, double y,
                          double *u, double *d)
{
   const int p = (int)x.size() - 1;
   if (p == 0) { u[0] = 1.0; d[0] = 0.0; return; }
   int k = 0;
   double lk = 1.0; // prod_{i != k} (y - x_i)
   for (; k < p; k++)
   {
      if (y >= (x[k] + x[k + 1]) / 2) { lk *= y - x[k]; }
      else
      {
         for (int i = k + 1; i <= p; i++) { lk *= y - x[i]; }
         break;
      }
   }
   const double l = lk * (y - x[k]);
   double sk = 0.0;
   for (int i = 0; i <= p; i++)
   {
      if (i == k) { continue; }
      const double si = 1.0 / (y - x[i]);
      sk += si;
      u[i] = l * si * bw[i];
   }
   u[k] = lk * bw[k];
   const double lp = l * sk + lk;
   for (int i = 0; i <= p; i++)
   {
      d[i] = (i == k) ? sk * u[k] : (lp * bw[i] - u[i]) / (y - x[i]);
   }
}

void h1_dof_to_quad(int p, int nq, const double *qpts, double *B, double *G)
{
   const int nd = p + 1;
   std::vector<double> x(nd), wts(nd), bw(nd, 1.0), u(nd), d(nd);
   gauss_lobatto(nd, x.data(), wts.data());
   for (int i = 0; i <= p; i++)
      for (int j = 0; j < i; j++)
      {
         const double xij = x[i] - x[j];
         bw[i] *= xij;
         bw[j] *= -xij;
      }
   for (int i = 0; i <= p; i++) { bw[i] = 1.0 / bw[i]; }
   for (int q = 0; q < nq; q++)
   {
      lagrange_eval(x, bw, qpts[q], u.data(), d.data());
      for (int j = 0; j < nd; j++)
      {
         B[q + nq * j] = u[j];
         G[q + nq * j] = d[j];
      }
   }
}

} // namespace b2f

extern "C" {

int b2f_gauss_legendre_h(int np, double *pts_h, double *wts_h)
{
   if (np < 1 || !pts_h || !wts_h) { b2f::set_error("b2f_gauss_legendre_h: bad arguments"); return B2F_ERR_ARG; }
   b2f::gauss_legendre(np, pts_h, wts_h);
   return B2F_OK;
}

int b2f_gauss_lobatto_h(int np, double *pts_h, double *wts_h)
{
   if (np < 1 || !pts_h || !wts_h) { b2f::set_error("b2f_gauss_lobatto_h: bad arguments"); return B2F_ERR_ARG; }
   b2f::gauss_lobatto(np, pts_h, wts_h);
   return B2F_OK;
}

int b2f_h1_dof_to_quad_h(int p, int nq, const double *qpts_h, double *B_h, double *G_h)
{
   if (p < 1 || nq < 1 || !qpts_h || !B_h || !G_h)
   {
      b2f::set_error("b2f_h1_dof_to_quad_h: bad arguments");
      return B2F_ERR_ARG;
   }
   b2f::h1_dof_to_quad(p, nq, qpts_h, B_h, G_h);
   return B2F_OK;
}

int b2f_tensor_weights_h(int dim, int nq, const double *w1, double *W)
{
   if ((dim != 2 && dim != 3) || nq < 1 || !w1 || !W)
   {
      b2f::set_error("b2f_tensor_weights_h: bad arguments");
      return B2F_ERR_ARG;
   }
   if (dim == 2)
   {
      for (int j = 0; j < nq; j++)
         for (int i = 0; i < nq; i++) { W[j * nq + i] = w1[i] * w1[j]; }
   }
   else
   {
      for (int k = 0; k < nq; k++)
         for (int j = 0; j < nq; j++)
            for (int i = 0; i < nq; i++) { W[(k * nq + j) * nq + i] = w1[i] * w1[j] * w1[k]; }
   }
   return B2F_OK;
}

int b2f_vector_randomize_h(double *v, int64_t n, int seed)
{
   if (n < 0 || (n > 0 && !v)) { b2f::set_error("b2f_vector_randomize_h: bad arguments"); return B2F_ERR_ARG; }
   const double scale = 1.0 / ((double)RAND_MAX + 1.0);
   srand(seed == 0 ? (unsigned)time(nullptr) : (unsigned)seed);
   for (int64_t i = 0; i < n; i++) { v[i] = std::fabs(rand() * scale); }
   return B2F_OK;
}

} // extern "C"
