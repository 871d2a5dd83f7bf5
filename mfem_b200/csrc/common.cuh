// mfem_b200/csrc/common.cuh -- shared declarations of libb200fem.so (internal).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200fem.h"

namespace b2f
{

void set_error(const char *fmt, ...);

#define B2F_CUDA(call)                                                                 \
   do                                                                                  \
   {                                                                                   \
      cudaError_t e_ = (call);                                                         \
      if (e_ != cudaSuccess)                                                           \
      {                                                                                \
         b2f::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,             \
                        cudaGetErrorString(e_));                                       \
         return B2F_ERR_CUDA;                                                          \
      }                                                                                \
   } while (0)

// kernel launch check, the reference does the same after every forall (forall.hpp:762)
#define B2F_LAUNCH_CHECK(ctx)                                                          \
   do                                                                                  \
   {                                                                                   \
      (ctx)->launches++;                                                               \
      cudaError_t e_ = cudaGetLastError();                                             \
      if (e_ != cudaSuccess)                                                           \
      {                                                                                \
         b2f::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,         \
                        cudaGetErrorString(e_));                                       \
         return B2F_ERR_CUDA;                                                          \
      }                                                                                \
   } while (0)

#define B2F_REQUIRE(cond, ...)                                                         \
   do                                                                                  \
   {                                                                                   \
      if (!(cond))                                                                     \
      {                                                                                \
         b2f::set_error(__VA_ARGS__);                                                  \
         return B2F_ERR_ARG;                                                           \
      }                                                                                \
   } while (0)

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

} // namespace b2f

struct b2f_ctx
{
   int device = 0;
   cudaStream_t stream = nullptr;
   int sm_count = 0;
   int max_smem_optin = 0;
   int64_t launches = 0;
   // scratch for reductions: device partials + device scalars + pinned host mirror
   double *d_partials = nullptr;
   int n_partials = 0;
   unsigned int *d_counter = nullptr;
   double *d_scalars = nullptr; // 64 doubles of device-resident scalars
   double *h_scalars = nullptr; // pinned
   // optional CUDA-event timing of the PA operator kernel launches (b2f_ctx_kernel_timing)
   int timing = 0;
   std::vector<cudaEvent_t> ev; // pairs (start, stop)
   int ev_used = 0;
};

namespace b2f
{
// Several contexts (devices) may live in one process: every entry point that allocates or launches
// makes its context's device current for the duration of the call and restores the caller's.
struct DeviceGuard
{
   int prev = -1;
   explicit DeviceGuard(const b2f_ctx *ctx)
   {
      if (!ctx) { return; }
      int cur = -1;
      if (cudaGetDevice(&cur) == cudaSuccess && cur != ctx->device)
      {
         prev = cur;
         cudaSetDevice(ctx->device);
      }
   }
   ~DeviceGuard() { if (prev >= 0) { cudaSetDevice(prev); } }
   DeviceGuard(const DeviceGuard &) = delete;
};
#define B2F_ON(ctx_expr) b2f::DeviceGuard b2f_device_guard_(ctx_expr)

// record an event on the context's stream when timing is enabled (pairs: before, after)
inline void timing_mark(b2f_ctx *ctx)
{
   if (!ctx->timing) { return; }
   if (ctx->ev_used == (int)ctx->ev.size())
   {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { return; }
      ctx->ev.push_back(e);
   }
   cudaEventRecord(ctx->ev[ctx->ev_used++], ctx->stream);
}
} // namespace b2f

namespace b2f
{
// opt-in shared memory and residency of a kernel, per (device, kernel): the attribute is per
// device and several contexts (devices, host threads) may share the process
template <class K>
inline int kernel_residency(b2f_ctx *ctx, K kern, int nt, size_t smem, int &resident)
{
   static std::map<std::pair<int, const void *>, int> cache;
   static std::mutex mtx;
   std::lock_guard<std::mutex> lock(mtx);
   int &r = cache[{ctx->device, (const void *)kern}];
   if (r == 0)
   {
      B2F_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int occ = 0;
      B2F_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem));
      r = (occ < 1 ? 1 : occ) * ctx->sm_count;
   }
   resident = r;
   return B2F_OK;
}

} // namespace b2f

struct b2f_restriction
{
   b2f_ctx *ctx;
   int ne, dof, ndofs, vdim, byvdim;
   int *gather_map = nullptr, *offsets = nullptr, *indices = nullptr;
};

enum { B2F_KIND_DIFFUSION = 0, B2F_KIND_MASS = 1 };

struct b2f_pa
{
   b2f_ctx *ctx;
   int kind, dim, d1d, q1d;
   int ncomp;      // quadrature-data components per point: 1 (mass), 3/4 (2D), 6/9 (3D)
   int symmetric;
   int64_t ne;
   double B[256], G[256]; // host copies, (q1d, d1d) column-major
   double *d_B = nullptr, *d_G = nullptr;
   double *data = nullptr; // (Q^dim, ncomp, NE)
   int64_t data_size = 0;
   int collocated = 0;     // B is the identity (BP5): interpolation is skipped
   int64_t rb = 0, re = -1; // element range of the next apply (re < 0: all), even-odd kernel only
   // vector (vdim > 1) form of the next apply, even-odd kernel only: component c of L-dof g is
   // x[g*vxs + c*vxo]; the gather index of component c starts at idx + c*vics
   int vd = 1;
   int64_t vxs = 1, vxo = 0, vics = 0;
   // |B|, |G|, |D| twin for AddAbsMultPA (built on first use, owned)
   b2f_pa *abs_twin = nullptr;
   // TMA tensor map over `data` (swizzled copies of the Q = 8 kernels, pa_apply_eo.cuh) and the
   // (data, ne) it was built for
   alignas(64) unsigned char tmap[128];
   const double *tmap_data = nullptr;
   int64_t tmap_ne = -1;
};

// A domain integrator restricted to the elements whose attribute is marked
// (BilinearForm::AddDomainIntegrator(bfi, elem_marker), bilinearform_ext.cpp:753-847).
// The enabled elements' quadrature data and gather-index rows are compacted at setup, so
// the fused kernels stream only those elements.
struct b2f_marked
{
   std::vector<unsigned char> on_h;  // [NE] element enabled
   std::vector<int64_t> prefix;      // [NE+1] enabled elements before e
   unsigned char *on = nullptr;      // device copy of on_h
   b2f_pa *cpa = nullptr;            // shallow copy of the integrator over the compacted data
   int *idx_c = nullptr;             // compacted gather_map rows
   int *fidx_c = nullptr;            // compacted fused_idx rows (vdim copies), if any
   int64_t ne_c = 0;
};

struct b2f_operator
{
   b2f_ctx *ctx;
   b2f_restriction *r;
   std::vector<b2f_pa *> integ;
   std::vector<b2f_marked *> marked;    // per integrator, nullptr = acts on every element
   double *te = nullptr;                // E-vector scratch of the marked unfused path
   int n_ess = 0;
   int *ess = nullptr;          // device list of essential ldofs / tdofs
   int *fused_idx = nullptr;    // gather_map with ess dofs marked -1 (constrained fused path);
                                // vdim copies (one per component) for vector spaces
   double *xe = nullptr, *ye = nullptr; // E-vectors for the unfused path (lazy)
   double *z = nullptr, *w = nullptr;   // L-vector scratch
   b2f_halo *halo = nullptr;
   double *xl = nullptr, *yl = nullptr; // L-vectors when a halo is attached
   std::vector<int> ess_l_h;            // essential LOCAL dofs as given at creation
   int signed_map = 0;                  // gather_map has negative (sign-flipped) entries
   int diag_policy = 1;                 // Operator::DiagonalPolicy: 0 = DIAG_ZERO, 1 = DIAG_ONE, 2 = DIAG_KEEP
   double *pcg_vec[4] = {nullptr, nullptr, nullptr, nullptr}; // r, d, z, q
   double *pcg_hist = nullptr;
   int pcg_hist_cap = 0;
   double *dot_acc = nullptr;           // 64 partial sums of the fused (d, A d)
   int64_t n_interior = -1;             // elements [0, n_interior) touch no neighbour-owned dof
   std::map<const void *, int64_t> scratch_cap; // capacity of each scratch vector above
};

namespace b2f
{
// pa_apply_*.cu : E-vector (idx == nullptr) or fused L-vector (idx = gather index, -1 = skip)
// dot_acc (64 doubles, or nullptr): also accumulate x . (A x) (fused form only; see
// pa_has_fused_dot)
int pa_apply(b2f_pa *pa, const double *x, double *y, const int *idx, double *dot_acc = nullptr);
bool pa_has_fused_dot(const b2f_pa *pa);
// pa_diag.cu: E-vector form (diag_e += diagonal), and fused with the unsigned scatter
// (diag_l[|gather_map[i]|] += diagonal, ElementRestriction::AbsMultTranspose)
int pa_diag(b2f_pa *pa, double *diag_e);
int pa_diag_fused(b2f_pa *pa, const int *gather_map, double *diag_l);
// pa_setup.cu
int diffusion_setup(b2f_ctx *ctx, int dim, int q1d, int64_t ne, const double *W_d,
                    const double *J, const double *coeff, int coeff_dim, int coeff_const,
                    double *D);
int mass_setup(b2f_ctx *ctx, int nq, int64_t ne, const double *W_d, const double *detJ,
               const double *coeff, int coeff_const, double *D);
int setup_from_nodes(b2f_ctx *ctx, int kind, int dim, int q1d, int64_t ne, const double *W_d,
                     int nd, const double *Bn_d, const double *Gn_d, const double *nodes_e,
                     double coeff, double *D);
// vector_ops.cu
int fill(b2f_ctx *ctx, double *x, double v, int64_t n);
// kernel-based zero fill / copy of the solver's hot path (never a copy-engine operation)
int zero_vector(b2f_ctx *ctx, double *x, int64_t n);
int copy_vector(b2f_ctx *ctx, double *dst, const double *src, int64_t n);
int set_entries(b2f_ctx *ctx, double *x, const int *idx, int n, double v);
int copy_entries(b2f_ctx *ctx, double *dst, const double *src, const int *idx, int n);
// dst[idx] = src[idx] and acc[0] += sum src[idx]^2
int copy_entries_sqsum(b2f_ctx *ctx, double *dst, const double *src, const int *idx, int n, double *acc);
// *out = sum acc[0..63]; acc[] = 0
int collect_dot(b2f_ctx *ctx, double *acc, double *out);
int dot_to_device(b2f_ctx *ctx, int64_t n, const double *x, const double *y, double *d_out);
// n device doubles -> pinned host memory through a store kernel (no copy engine), then a stream sync
int publish_scalars(b2f_ctx *ctx, const double *d_src, double *h_dst, int n);
} // namespace b2f
