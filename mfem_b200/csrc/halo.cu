// mfem_b200/csrc/halo.cu -- rank-parallel conforming prolongation P / P^T and the
// all-reduced inner product.
//
// Replaces ConformingProlongationOperator / DeviceConformingProlongationOperator
// (fem/pfespace.hpp:615-697, fem/pfespace.cpp:5479-5927) and InnerProduct(comm, x, y)
// (linalg/vector.hpp:773-779).  The reference packs shared dofs with a device kernel,
// stages through the host (unless GPU-aware MPI) and exchanges with MPI_Isend/Irecv per
// neighbour; here the packed buffers go device-to-device over NVLink with one grouped
// ncclSend/ncclRecv per neighbour, enqueued on the context's stream (no host sync), and
// the PCG scalars are ncclAllReduce'd in place in device memory.
//
// Semantics kept from the reference:
//   P   : y_l[ltdof_ldof[i]] = x_t[i];  owners send x_t[shr_ltdof[k]];  sharers store the
//         received values at y_l[ext_ldof[k]]                      (pfespace.cpp:5760-5816)
//   P^T : y_t[i] = x_l[ltdof_ldof[i]];  sharers send x_l[ext_ldof[k]];  owners add the
//         received values per true dof in a FIXED order (the CSR unq_shr_i/unq_shr_j of
//         pfespace.cpp:5833-5860: ascending position in the receive buffer), so the
//         result is deterministic.
// NCCL is resolved at run time (dlopen) so that single-rank use has no NCCL dependency
// and the copy already loaded by the host process (e.g. torch's) is shared.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>

#include "common.cuh"

namespace b2f
{

struct NcclApi
{
   void *lib = nullptr;
   ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
   ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
   ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
   ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                             cudaStream_t) = nullptr;
   ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*GroupStart)() = nullptr;
   ncclResult_t (*GroupEnd)() = nullptr;
   const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load()
{
   if (g_nccl.lib) { return B2F_OK; }
   const char *names[] = {getenv("B2F_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
   void *lib = nullptr;
   for (const char *n : names)
   {
      if (n && *n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) { break; }
   }
   if (!lib)
   {
      set_error("NCCL not found (tried $B2F_NCCL_LIB, libnccl.so.2, libnccl.so): %s", dlerror());
      return B2F_ERR_NCCL;
   }
#define B2F_SYM(name)                                                                  \
   *(void **)(&g_nccl.name) = dlsym(lib, "nccl" #name);                                \
   if (!g_nccl.name)                                                                   \
   {                                                                                   \
      set_error("NCCL symbol nccl" #name " missing");                                  \
      return B2F_ERR_NCCL;                                                             \
   }
   B2F_SYM(GetUniqueId) B2F_SYM(CommInitRank) B2F_SYM(CommDestroy) B2F_SYM(AllReduce)
   B2F_SYM(Send) B2F_SYM(Recv) B2F_SYM(GroupStart) B2F_SYM(GroupEnd) B2F_SYM(GetErrorString)
#undef B2F_SYM
   g_nccl.lib = lib;
   return B2F_OK;
}

#define B2F_NCCL(call)                                                                 \
   do                                                                                  \
   {                                                                                   \
      ncclResult_t r_ = (call);                                                        \
      if (r_ != ncclSuccess)                                                           \
      {                                                                                \
         set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                  \
                   g_nccl.GetErrorString(r_));                                         \
         return B2F_ERR_NCCL;                                                          \
      }                                                                                \
   } while (0)

} // namespace b2f

struct b2f_halo
{
   b2f_ctx *ctx = nullptr;
   int rank = 0, nranks = 1;
   ncclComm_t comm = nullptr;
   int n_ldofs = 0, n_tdofs = 0, n_nbr = 0;
   std::vector<int> nbr_rank, shr_off, ext_off;
   std::vector<int> l2t_h;        // [n_ldofs] tdof of an owned ldof, else -1
   int *ltdof_ldof = nullptr;     // [n_tdofs]
   int *shr_ltdof = nullptr;      // [n_shr]
   int *ext_ldof = nullptr;       // [n_ext]
   int n_shr = 0, n_ext = 0;
   int n_unq = 0;                 // unique shared true dofs
   int *unq_ltdof = nullptr;      // [n_unq]
   int *unq_i = nullptr;          // [n_unq+1] CSR into unq_j
   int *unq_j = nullptr;          // [n_shr] positions in shr_buf, ascending per row
   double *shr_buf = nullptr, *ext_buf = nullptr;
   bool prefix = false;           // ltdof_ldof is the identity: a T-vector is the head of an L-vector
   // ---- peer-memory exchange (b2f_halo_p2p_connect): the pack kernel stores straight into
   // the neighbour's receive buffer over NVLink and publishes an epoch flag; no NCCL call
   bool p2p = false;
   unsigned long long *flags = nullptr;      // [64] local: [k] P epoch of neighbour k, [32+k] P^T epoch
   unsigned int *tickets = nullptr;          // [64] local block tickets of the push kernels
   int *shr_off_d = nullptr, *ext_off_d = nullptr;
   double **peer_ext_dst = nullptr;          // [n_nbr] where my shared values land in neighbour k
   double **peer_shr_dst = nullptr;          // [n_nbr] where my contributions land in neighbour k
   unsigned long long **peer_flag_p = nullptr, **peer_flag_pt = nullptr;
   unsigned long long epoch_p = 0, epoch_pt = 0;
   // Protocol invariant of the peer-memory exchange: on every rank a P push is followed by exactly
   // one P^T push before the next P push (one operator application = P ... P^T).  That alternation
   // is what makes the single receive buffer per direction race-free: a neighbour can only start
   // its next P push after it has waited for my P^T data, which I send after consuming its P data.
   // phase: 0 = a P push is next, 1 = a P^T push is next.  A call out of order (an application that
   // failed half-way) is refused here instead of leaving the neighbours spinning on a flag.
   int phase = 0;
   int max_shr = 0, max_ext = 0;
   std::vector<void *> opened;               // cudaIpcOpenMemHandle results to close
};

// what a rank publishes so that its neighbours can address its receive buffers
struct b2f_p2p_blob
{
   int rank, n_nbr;
   int nbr_rank[32];
   int shr_off[33], ext_off[33];
   cudaIpcMemHandle_t ext_buf, shr_buf, flags;
};

namespace b2f
{

// y[dst[i]] = x[i]
__global__ void __launch_bounds__(256)
scatter_to_kernel(int n, const int *__restrict__ dst, const double *__restrict__ x, double *y)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { y[__ldg(dst + i)] = x[i]; }
}
// y[i] = x[src[i]]
__global__ void __launch_bounds__(256)
gather_from_kernel(int n, const int *__restrict__ src, const double *__restrict__ x, double *y)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { y[i] = __ldg(x + __ldg(src + i)); }
}
// y[unq_ltdof[i]] += sum_{j in row i} buf[unq_j[j]]   (fixed order)
__global__ void __launch_bounds__(256)
assemble_shared_kernel(int n, const int *__restrict__ unq_ltdof, const int *__restrict__ unq_i,
                       const int *__restrict__ unq_j, const double *__restrict__ buf, double *y)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   const int t = unq_ltdof[i];
   double v = y[t];
   for (int j = unq_i[i]; j < unq_i[i + 1]; ++j) { v += buf[unq_j[j]]; }
   y[t] = v;
}

// ---- peer-memory exchange -----------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
   asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
   unsigned long long v;
   asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}

// Fused pack + send: block column blockIdx.y serves neighbour k and stores
// x[src_idx[off[k] + i]] straight into that neighbour's receive buffer (a peer pointer mapped
// over NVLink).  The last block of a column to finish publishes `epoch` in the neighbour's flag.
__global__ void __launch_bounds__(256)
p2p_push_kernel(const int *__restrict__ src_idx, const int *__restrict__ off,
                const double *__restrict__ x, double *const *__restrict__ peer_dst,
                unsigned long long *const *__restrict__ peer_flag, unsigned int *tickets,
                unsigned long long epoch)
{
   const int k = blockIdx.y;
   const int b = off[k], n = off[k + 1] - b;
   double *dst = peer_dst[k];
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      dst[i] = __ldg(x + __ldg(src_idx + b + i));
   }
   __threadfence_system();
   __syncthreads();
   if (threadIdx.x == 0)
   {
      const unsigned int t = atomicAdd(tickets + k, 1u);
      if (t == gridDim.x - 1)
      {
         tickets[k] = 0;
         __threadfence_system();
         st_release_sys(peer_flag[k], epoch);
      }
   }
}

// every block first waits until all neighbours have published `epoch`
__device__ __forceinline__ void p2p_wait(const unsigned long long *flags, int n_nbr,
                                         unsigned long long epoch)
{
   if (threadIdx.x < n_nbr)
   {
      // bounded: a neighbour that never publishes (crashed rank, mismatched call sequence) must
      // become a CUDA error on this rank, not a hung GPU (~20 s at 256 ns per poll)
      unsigned int polls = 0;
      while (ld_acquire_sys(flags + threadIdx.x) < epoch)
      {
         __nanosleep(256);
         if (++polls > 80000000u)
         {
            printf("b200fem halo: neighbour %d did not publish epoch %llu\n", (int)threadIdx.x, epoch);
            __trap();
         }
      }
   }
   __syncthreads();
}

// P, receiving side: wait, then y[ext_ldof[i]] = ext_buf[i]
__global__ void __launch_bounds__(256)
p2p_wait_scatter_kernel(int n, const int *__restrict__ dst, const double *x, double *y,
                        const unsigned long long *flags, int n_nbr, unsigned long long epoch)
{
   p2p_wait(flags, n_nbr, epoch);
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { y[__ldg(dst + i)] = x[i]; }
}

// P^T, receiving side: wait, then the fixed-order assembly of assemble_shared_kernel
__global__ void __launch_bounds__(256)
p2p_wait_assemble_kernel(int n, const int *__restrict__ unq_ltdof, const int *__restrict__ unq_i,
                         const int *__restrict__ unq_j, const double *buf, double *y,
                         const unsigned long long *flags, int n_nbr, unsigned long long epoch)
{
   p2p_wait(flags, n_nbr, epoch);
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   const int t = unq_ltdof[i];
   double v = y[t];
   for (int j = unq_i[i]; j < unq_i[i + 1]; ++j) { v += buf[unq_j[j]]; }
   y[t] = v;
}

static int p2p_p_push(b2f_halo *h, const double *v)
{
   b2f_ctx *ctx = h->ctx;
   if (h->n_nbr == 0) { return B2F_OK; }
   if (h->phase != 0)
   {
      set_error("halo (peer-memory): P exchange started while a P^T exchange is pending -- the ranks are out of step");
      return B2F_ERR_ARG;
   }
   const unsigned long long ep = h->epoch_p + 1;
   const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(h->max_shr, 256), 64)), h->n_nbr);
   p2p_push_kernel<<<grid, 256, 0, ctx->stream>>>(h->shr_ltdof, h->shr_off_d, v, h->peer_ext_dst,
                                                  h->peer_flag_p, h->tickets, ep);
   B2F_LAUNCH_CHECK(ctx);
   h->epoch_p = ep; // only a launch that went out advances the epoch
   h->phase = 1;
   return B2F_OK;
}

static int p2p_p_wait(b2f_halo *h, double *v)
{
   b2f_ctx *ctx = h->ctx;
   cudaStream_t s = ctx->stream;
   if (h->n_nbr == 0) { return B2F_OK; }
   const unsigned long long ep = h->epoch_p;
   p2p_wait_scatter_kernel<<<(unsigned)std::max<int64_t>(1, cdiv(h->n_ext, 256)), 256, 0, s>>>(
      h->n_ext, h->ext_ldof, h->ext_buf, v, h->flags, h->n_nbr, ep);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

static int p2p_mult_inplace(b2f_halo *h, double *v)
{
   int rc = p2p_p_push(h, v);
   return rc ? rc : p2p_p_wait(h, v);
}

static int p2p_pt_push(b2f_halo *h, const double *v)
{
   b2f_ctx *ctx = h->ctx;
   if (h->n_nbr == 0) { return B2F_OK; }
   if (h->phase != 1)
   {
      set_error("halo (peer-memory): P^T exchange without a preceding P exchange -- the ranks are out of step");
      return B2F_ERR_ARG;
   }
   const unsigned long long ep = h->epoch_pt + 1;
   const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(h->max_ext, 256), 64)), h->n_nbr);
   p2p_push_kernel<<<grid, 256, 0, ctx->stream>>>(h->ext_ldof, h->ext_off_d, v, h->peer_shr_dst,
                                                  h->peer_flag_pt, h->tickets + 32, ep);
   B2F_LAUNCH_CHECK(ctx);
   h->epoch_pt = ep;
   h->phase = 0;
   return B2F_OK;
}

static int p2p_pt_wait(b2f_halo *h, double *v)
{
   b2f_ctx *ctx = h->ctx;
   cudaStream_t s = ctx->stream;
   if (h->n_nbr == 0) { return B2F_OK; }
   const unsigned long long ep = h->epoch_pt;
   p2p_wait_assemble_kernel<<<(unsigned)std::max<int64_t>(1, cdiv(h->n_unq, 256)), 256, 0, s>>>(
      h->n_unq, h->unq_ltdof, h->unq_i, h->unq_j, h->shr_buf, v, h->flags + 32, h->n_nbr, ep);
   B2F_LAUNCH_CHECK(ctx);
   return B2F_OK;
}

static int p2p_mult_transpose_inplace(b2f_halo *h, double *v)
{
   int rc = p2p_pt_push(h, v);
   return rc ? rc : p2p_pt_wait(h, v);
}

// split-phase forms for overlapping the exchange with element work (peer-memory mode only)
bool halo_is_p2p(const b2f_halo *h) { return h->p2p; }
int halo_p_push(b2f_halo *h, const double *v) { return p2p_p_push(h, v); }
int halo_p_wait(b2f_halo *h, double *v) { return p2p_p_wait(h, v); }
int halo_pt_push(b2f_halo *h, const double *v) { return p2p_pt_push(h, v); }
int halo_pt_wait(b2f_halo *h, double *v) { return p2p_pt_wait(h, v); }

static int exchange(b2f_halo *h, const double *send, const std::vector<int> &send_off,
                    double *recv, const std::vector<int> &recv_off)
{
   if (h->n_nbr == 0) { return B2F_OK; }
   cudaStream_t s = h->ctx->stream;
   B2F_NCCL(g_nccl.GroupStart());
   for (int k = 0; k < h->n_nbr; k++)
   {
      const int ns = send_off[k + 1] - send_off[k], nr = recv_off[k + 1] - recv_off[k];
      if (ns > 0) { B2F_NCCL(g_nccl.Send(send + send_off[k], ns, ncclDouble, h->nbr_rank[k], h->comm, s)); }
      if (nr > 0) { B2F_NCCL(g_nccl.Recv(recv + recv_off[k], nr, ncclDouble, h->nbr_rank[k], h->comm, s)); }
   }
   B2F_NCCL(g_nccl.GroupEnd());
   return B2F_OK;
}

int halo_mult(b2f_halo *h, const double *x_t, double *y_l)
{
   b2f_ctx *ctx = h->ctx;
   cudaStream_t s = ctx->stream;
   if (h->n_shr)
   {
      gather_from_kernel<<<(unsigned)cdiv(h->n_shr, 256), 256, 0, s>>>(h->n_shr, h->shr_ltdof, x_t,
                                                                        h->shr_buf);
      B2F_LAUNCH_CHECK(ctx);
   }
   int rc = exchange(h, h->shr_buf, h->shr_off, h->ext_buf, h->ext_off);
   if (rc) { return rc; }
   if (h->n_tdofs)
   {
      scatter_to_kernel<<<(unsigned)cdiv(h->n_tdofs, 256), 256, 0, s>>>(h->n_tdofs, h->ltdof_ldof,
                                                                         x_t, y_l);
      B2F_LAUNCH_CHECK(ctx);
   }
   if (h->n_ext)
   {
      scatter_to_kernel<<<(unsigned)cdiv(h->n_ext, 256), 256, 0, s>>>(h->n_ext, h->ext_ldof,
                                                                       h->ext_buf, y_l);
      B2F_LAUNCH_CHECK(ctx);
   }
   return B2F_OK;
}

int halo_mult_transpose(b2f_halo *h, const double *x_l, double *y_t)
{
   b2f_ctx *ctx = h->ctx;
   cudaStream_t s = ctx->stream;
   if (h->n_ext)
   {
      gather_from_kernel<<<(unsigned)cdiv(h->n_ext, 256), 256, 0, s>>>(h->n_ext, h->ext_ldof, x_l,
                                                                        h->ext_buf);
      B2F_LAUNCH_CHECK(ctx);
   }
   int rc = exchange(h, h->ext_buf, h->ext_off, h->shr_buf, h->shr_off);
   if (rc) { return rc; }
   if (h->n_tdofs)
   {
      gather_from_kernel<<<(unsigned)cdiv(h->n_tdofs, 256), 256, 0, s>>>(h->n_tdofs, h->ltdof_ldof,
                                                                          x_l, y_t);
      B2F_LAUNCH_CHECK(ctx);
   }
   if (h->n_unq)
   {
      assemble_shared_kernel<<<(unsigned)cdiv(h->n_unq, 256), 256, 0, s>>>(
         h->n_unq, h->unq_ltdof, h->unq_i, h->unq_j, h->shr_buf, y_t);
      B2F_LAUNCH_CHECK(ctx);
   }
   return B2F_OK;
}

// In-place forms for the owned-first numbering (ltdof_ldof = identity): v has n_ldofs entries
// whose head is the true-dof vector.  P only fills the tail, P^T only adds into the head.
bool halo_is_prefix(const b2f_halo *h) { return h->prefix; }

int halo_mult_inplace(b2f_halo *h, double *v)
{
   if (h->p2p) { return p2p_mult_inplace(h, v); }
   b2f_ctx *ctx = h->ctx;
   cudaStream_t s = ctx->stream;
   if (h->n_shr)
   {
      gather_from_kernel<<<(unsigned)cdiv(h->n_shr, 256), 256, 0, s>>>(h->n_shr, h->shr_ltdof, v, h->shr_buf);
      B2F_LAUNCH_CHECK(ctx);
   }
   int rc = exchange(h, h->shr_buf, h->shr_off, h->ext_buf, h->ext_off);
   if (rc) { return rc; }
   if (h->n_ext)
   {
      scatter_to_kernel<<<(unsigned)cdiv(h->n_ext, 256), 256, 0, s>>>(h->n_ext, h->ext_ldof, h->ext_buf, v);
      B2F_LAUNCH_CHECK(ctx);
   }
   return B2F_OK;
}

int halo_mult_transpose_inplace(b2f_halo *h, double *v)
{
   if (h->p2p) { return p2p_mult_transpose_inplace(h, v); }
   b2f_ctx *ctx = h->ctx;
   cudaStream_t s = ctx->stream;
   if (h->n_ext)
   {
      gather_from_kernel<<<(unsigned)cdiv(h->n_ext, 256), 256, 0, s>>>(h->n_ext, h->ext_ldof, v, h->ext_buf);
      B2F_LAUNCH_CHECK(ctx);
   }
   int rc = exchange(h, h->ext_buf, h->ext_off, h->shr_buf, h->shr_off);
   if (rc) { return rc; }
   if (h->n_unq)
   {
      assemble_shared_kernel<<<(unsigned)cdiv(h->n_unq, 256), 256, 0, s>>>(
         h->n_unq, h->unq_ltdof, h->unq_i, h->unq_j, h->shr_buf, v);
      B2F_LAUNCH_CHECK(ctx);
   }
   return B2F_OK;
}

int halo_allreduce_sum(b2f_halo *h, double *d_scalar, int count)
{
   if (h->nranks == 1) { return B2F_OK; }
   B2F_NCCL(g_nccl.AllReduce(d_scalar, d_scalar, count, ncclDouble, ncclSum, h->comm,
                             h->ctx->stream));
   return B2F_OK;
}

int halo_sizes(const b2f_halo *h, int *n_ldofs, int *n_tdofs)
{
   *n_ldofs = h->n_ldofs;
   *n_tdofs = h->n_tdofs;
   return B2F_OK;
}

const int *halo_ldof_to_tdof_h(const b2f_halo *h) { return h->l2t_h.data(); }

template <typename T>
static int up(const std::vector<T> &v, T **d)
{
   *d = nullptr;
   B2F_CUDA(cudaMalloc((void **)d, sizeof(T) * (v.size() ? v.size() : 1)));
   if (!v.empty()) { B2F_CUDA(cudaMemcpy(*d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice)); }
   return B2F_OK;
}

} // namespace b2f

using namespace b2f;

extern "C" {

int b2f_nccl_unique_id_h(void *id128_h)
{
   B2F_REQUIRE(id128_h, "b2f_nccl_unique_id_h: NULL argument");
   static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
   int rc = nccl_load();
   if (rc) { return rc; }
   ncclUniqueId id;
   B2F_NCCL(g_nccl.GetUniqueId(&id));
   memcpy(id128_h, &id, sizeof(id));
   return B2F_OK;
}

int b2f_halo_create(b2f_ctx *ctx, const void *nccl_id128_h, int rank, int nranks, int n_ldofs,
                    int n_tdofs, const int *ltdof_ldof_h, int n_nbr, const int *nbr_rank_h,
                    const int *shr_offsets_h, const int *shr_ltdof_h, const int *ext_offsets_h,
                    const int *ext_ldof_h, b2f_halo **out)
{
   B2F_REQUIRE(ctx && out, "b2f_halo_create: NULL argument");
   B2F_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "b2f_halo_create: rank %d of %d", rank, nranks);
   B2F_REQUIRE(n_ldofs >= 0 && n_tdofs >= 0 && n_tdofs <= n_ldofs && (ltdof_ldof_h || !n_tdofs),
               "b2f_halo_create: bad dof counts (%d ldofs, %d tdofs)", n_ldofs, n_tdofs);
   B2F_REQUIRE(n_nbr >= 0 && (n_nbr == 0 || (nbr_rank_h && shr_offsets_h && ext_offsets_h)),
               "b2f_halo_create: neighbour tables are NULL");
   B2F_REQUIRE(nranks > 1 || n_nbr == 0, "b2f_halo_create: neighbours given for a single rank");
   B2F_REQUIRE(nranks == 1 || nccl_id128_h, "b2f_halo_create: nccl id is NULL");
   b2f_halo *h = new b2f_halo;
   h->ctx = ctx; h->rank = rank; h->nranks = nranks;
   h->n_ldofs = n_ldofs; h->n_tdofs = n_tdofs; h->n_nbr = n_nbr;
   h->nbr_rank.assign(nbr_rank_h, nbr_rank_h + n_nbr);
   h->shr_off.assign(1, 0); h->ext_off.assign(1, 0);
   if (n_nbr)
   {
      h->shr_off.assign(shr_offsets_h, shr_offsets_h + n_nbr + 1);
      h->ext_off.assign(ext_offsets_h, ext_offsets_h + n_nbr + 1);
   }
   h->n_shr = h->shr_off.back(); h->n_ext = h->ext_off.back();
   auto fail = [&](const char *msg, int v) { set_error(msg, v); b2f_halo_destroy(h); return B2F_ERR_ARG; };
   for (int k = 0; k < n_nbr; k++)
   {
      if (nbr_rank_h[k] < 0 || nbr_rank_h[k] >= nranks || nbr_rank_h[k] == rank)
      { return fail("b2f_halo_create: bad neighbour rank %d", nbr_rank_h[k]); }
   }
   h->l2t_h.assign(n_ldofs, -1);
   h->prefix = true;
   for (int i = 0; i < n_tdofs; i++)
   {
      h->prefix = h->prefix && ltdof_ldof_h[i] == i;
      const int l = ltdof_ldof_h[i];
      if (l < 0 || l >= n_ldofs || h->l2t_h[l] != -1) { return fail("b2f_halo_create: bad ltdof_ldof entry %d", l); }
      h->l2t_h[l] = i;
   }
   for (int k = 0; k < h->n_shr; k++)
   {
      if (shr_ltdof_h[k] < 0 || shr_ltdof_h[k] >= n_tdofs) { return fail("b2f_halo_create: bad shr_ltdof entry %d", shr_ltdof_h[k]); }
   }
   for (int k = 0; k < h->n_ext; k++)
   {
      const int l = ext_ldof_h[k];
      if (l < 0 || l >= n_ldofs || h->l2t_h[l] != -1) { return fail("b2f_halo_create: ext ldof %d is out of range or owned", l); }
   }
   // unique shared true dofs with their receive-buffer positions (pfespace.cpp:5833-5860)
   std::vector<std::pair<int, int>> pairs(h->n_shr);
   for (int k = 0; k < h->n_shr; k++) { pairs[k] = {shr_ltdof_h[k], k}; }
   std::sort(pairs.begin(), pairs.end());
   std::vector<int> unq, ui(1, 0), uj(h->n_shr);
   for (int k = 0; k < h->n_shr; k++)
   {
      if (k == 0 || pairs[k].first != pairs[k - 1].first)
      {
         if (k) { ui.push_back(k); }
         unq.push_back(pairs[k].first);
      }
      uj[k] = pairs[k].second;
   }
   if (h->n_shr) { ui.push_back(h->n_shr); }
   h->n_unq = (int)unq.size();
   std::vector<int> v_ltdof(ltdof_ldof_h, ltdof_ldof_h + n_tdofs);
   std::vector<int> v_shr(shr_ltdof_h, shr_ltdof_h + h->n_shr), v_ext(ext_ldof_h, ext_ldof_h + h->n_ext);
   int rc = B2F_OK;
   if (!rc) { rc = up(v_ltdof, &h->ltdof_ldof); }
   if (!rc) { rc = up(v_shr, &h->shr_ltdof); }
   if (!rc) { rc = up(v_ext, &h->ext_ldof); }
   if (!rc) { rc = up(unq, &h->unq_ltdof); }
   if (!rc) { rc = up(ui, &h->unq_i); }
   if (!rc) { rc = up(uj, &h->unq_j); }
   if (rc) { b2f_halo_destroy(h); return rc; }
   if (cudaMalloc((void **)&h->shr_buf, sizeof(double) * (h->n_shr ? h->n_shr : 1)) != cudaSuccess ||
       cudaMalloc((void **)&h->ext_buf, sizeof(double) * (h->n_ext ? h->n_ext : 1)) != cudaSuccess)
   {
      set_error("b2f_halo_create: out of device memory");
      b2f_halo_destroy(h);
      return B2F_ERR_CUDA;
   }
   if (nranks > 1)
   {
      rc = nccl_load();
      if (rc) { b2f_halo_destroy(h); return rc; }
      ncclUniqueId id;
      memcpy(&id, nccl_id128_h, sizeof(id));
      ncclResult_t r = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
      if (r != ncclSuccess)
      {
         set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
         h->comm = nullptr;
         b2f_halo_destroy(h);
         return B2F_ERR_NCCL;
      }
   }
   *out = h;
   return B2F_OK;
}

int b2f_halo_p2p_blob_bytes(void) { return (int)sizeof(b2f_p2p_blob); }

int b2f_halo_p2p_export(b2f_halo *h, void *blob_h)
{
   B2F_REQUIRE(h && blob_h, "b2f_halo_p2p_export: NULL argument");
   B2F_REQUIRE(h->n_nbr <= 32, "b2f_halo_p2p_export: more than 32 neighbours");
   b2f_p2p_blob b;
   memset(&b, 0, sizeof(b));
   b.rank = h->rank; b.n_nbr = h->n_nbr;
   for (int k = 0; k < h->n_nbr; k++) { b.nbr_rank[k] = h->nbr_rank[k]; }
   for (int k = 0; k <= h->n_nbr; k++) { b.shr_off[k] = h->shr_off[k]; b.ext_off[k] = h->ext_off[k]; }
   if (!h->flags)
   {
      B2F_CUDA(cudaMalloc((void **)&h->flags, sizeof(unsigned long long) * 64));
      B2F_CUDA(cudaMemset(h->flags, 0, sizeof(unsigned long long) * 64));
      B2F_CUDA(cudaMalloc((void **)&h->tickets, sizeof(unsigned int) * 64));
      B2F_CUDA(cudaMemset(h->tickets, 0, sizeof(unsigned int) * 64));
   }
   B2F_CUDA(cudaIpcGetMemHandle(&b.ext_buf, h->ext_buf));
   B2F_CUDA(cudaIpcGetMemHandle(&b.shr_buf, h->shr_buf));
   B2F_CUDA(cudaIpcGetMemHandle(&b.flags, h->flags));
   memcpy(blob_h, &b, sizeof(b));
   return B2F_OK;
}

int b2f_halo_p2p_connect(b2f_halo *h, const void *blobs_h, int nranks)
{
   B2F_REQUIRE(h && blobs_h && nranks == h->nranks, "b2f_halo_p2p_connect: bad arguments");
   B2F_REQUIRE(h->prefix, "b2f_halo_p2p_connect: needs the owned-first local numbering");
   B2F_REQUIRE(h->flags, "b2f_halo_p2p_connect: call b2f_halo_p2p_export first");
   const b2f_p2p_blob *blobs = (const b2f_p2p_blob *)blobs_h;
   std::vector<double *> ext_dst(h->n_nbr), shr_dst(h->n_nbr);
   std::vector<unsigned long long *> fp(h->n_nbr), fpt(h->n_nbr);
   h->max_shr = h->max_ext = 0;
   for (int k = 0; k < h->n_nbr; k++)
   {
      const b2f_p2p_blob &nb = blobs[h->nbr_rank[k]];
      B2F_REQUIRE(nb.rank == h->nbr_rank[k], "b2f_halo_p2p_connect: blob %d is from rank %d", h->nbr_rank[k], nb.rank);
      int j = -1;
      for (int q = 0; q < nb.n_nbr; q++) { if (nb.nbr_rank[q] == h->rank) { j = q; } }
      B2F_REQUIRE(j >= 0, "b2f_halo_p2p_connect: rank %d does not list rank %d as a neighbour", nb.rank, h->rank);
      const int my_shr = h->shr_off[k + 1] - h->shr_off[k], my_ext = h->ext_off[k + 1] - h->ext_off[k];
      B2F_REQUIRE(my_shr == nb.ext_off[j + 1] - nb.ext_off[j] && my_ext == nb.shr_off[j + 1] - nb.shr_off[j],
                  "b2f_halo_p2p_connect: message sizes with rank %d do not match", nb.rank);
      void *pe = nullptr, *ps = nullptr, *pf = nullptr;
      B2F_CUDA(cudaIpcOpenMemHandle(&pe, nb.ext_buf, cudaIpcMemLazyEnablePeerAccess));
      h->opened.push_back(pe);
      B2F_CUDA(cudaIpcOpenMemHandle(&ps, nb.shr_buf, cudaIpcMemLazyEnablePeerAccess));
      h->opened.push_back(ps);
      B2F_CUDA(cudaIpcOpenMemHandle(&pf, nb.flags, cudaIpcMemLazyEnablePeerAccess));
      h->opened.push_back(pf);
      ext_dst[k] = (double *)pe + nb.ext_off[j];
      shr_dst[k] = (double *)ps + nb.shr_off[j];
      fp[k] = (unsigned long long *)pf + j;
      fpt[k] = (unsigned long long *)pf + 32 + j;
      h->max_shr = std::max(h->max_shr, my_shr);
      h->max_ext = std::max(h->max_ext, my_ext);
   }
   int rc = B2F_OK;
   if (!rc) { rc = up(h->shr_off, &h->shr_off_d); }
   if (!rc) { rc = up(h->ext_off, &h->ext_off_d); }
   if (!rc) { rc = up(ext_dst, &h->peer_ext_dst); }
   if (!rc) { rc = up(shr_dst, &h->peer_shr_dst); }
   if (!rc) { rc = up(fp, &h->peer_flag_p); }
   if (!rc) { rc = up(fpt, &h->peer_flag_pt); }
   if (rc) { return rc; }
   h->p2p = true;
   return B2F_OK;
}

void b2f_halo_destroy(b2f_halo *h)
{
   if (!h) { return; }
   cudaStreamSynchronize(h->ctx->stream);
   for (void *p : h->opened) { cudaIpcCloseMemHandle(p); }
   cudaFree(h->flags); cudaFree(h->tickets); cudaFree(h->shr_off_d); cudaFree(h->ext_off_d);
   cudaFree(h->peer_ext_dst); cudaFree(h->peer_shr_dst); cudaFree(h->peer_flag_p); cudaFree(h->peer_flag_pt);
   if (h->comm) { cudaStreamSynchronize(h->ctx->stream); g_nccl.CommDestroy(h->comm); }
   cudaFree(h->ltdof_ldof); cudaFree(h->shr_ltdof); cudaFree(h->ext_ldof);
   cudaFree(h->unq_ltdof); cudaFree(h->unq_i); cudaFree(h->unq_j);
   cudaFree(h->shr_buf); cudaFree(h->ext_buf);
   delete h;
}

int b2f_halo_mult(b2f_halo *h, const double *x_t, double *y_l)
{
   B2F_REQUIRE(h && (x_t || !h->n_tdofs) && (y_l || !h->n_ldofs), "b2f_halo_mult: NULL argument");
   return halo_mult(h, x_t, y_l);
}

int b2f_halo_mult_transpose(b2f_halo *h, const double *x_l, double *y_t)
{
   B2F_REQUIRE(h && (x_l || !h->n_ldofs) && (y_t || !h->n_tdofs),
               "b2f_halo_mult_transpose: NULL argument");
   return halo_mult_transpose(h, x_l, y_t);
}

int b2f_halo_dot(b2f_halo *h, int64_t n, const double *x, const double *y, double *result_h)
{
   B2F_REQUIRE(h && result_h && (n == 0 || (x && y)), "b2f_halo_dot: NULL argument");
   b2f_ctx *ctx = h->ctx;
   double *slot = ctx->d_scalars + 62;
   if (n > 0) { int rc = dot_to_device(ctx, n, x, y, slot); if (rc) { return rc; } }
   else { B2F_CUDA(cudaMemsetAsync(slot, 0, sizeof(double), ctx->stream)); }
   int rc = halo_allreduce_sum(h, slot, 1);
   if (rc) { return rc; }
   rc = publish_scalars(ctx, slot, ctx->h_scalars + 62, 1);
   if (rc) { return rc; }
   *result_h = ctx->h_scalars[62];
   return B2F_OK;
}

} // extern "C"
