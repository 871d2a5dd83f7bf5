// mfem_b200/csrc/host_partition.cpp -- Cartesian block partition of the structured mesh and
// the shared-dof tables of the rank-parallel H1 space (host C++, integer work only).
//
// What it stands in for in the reference (which needs MPI + METIS to do the same):
//   Mesh::CartesianPartitioning                       mesh/mesh.cpp:9188
//   ParMesh(comm, mesh, partitioning)                 mesh/pmesh.cpp:106
//   GroupTopology::Create, master = lowest rank       general/communication.cpp:97-124
//   ParFiniteElementSpace::ConstructTrueDofs          fem/pfespace.cpp:2169
//     (a true dof is a local dof whose group master is this rank; ltdof order = ldof order;
//      here the local dofs are additionally renumbered owned-first, see below)
//   GroupCommunicator::GetNeighborLTDofTable / GetNeighborLDofTable
//                                                     general/communication.cpp:543,579
// Every rank runs this independently (no communication): a dof is identified by its point
// on the global (p*nx+1) x (p*ny+1) x (p*nz+1) lattice, the ranks that share it follow from
// the block extents, and both sides of every neighbour pair list their common dofs in
// ascending lattice order, so send and receive buffers line up.
#include <algorithm>
#include <exception>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <map>
#include <thread>
#include <vector>

#include "../../include/b200fem.h"
#include "host_parallel.h"

namespace b2f
{
void set_error(const char *fmt, ...);
}

namespace
{
struct Split
{
   int n, parts;
   int lo(int b) const { return (int)((int64_t)b * n / parts); }
};

// The caller asks twice (sizes, then the arrays): the tables of the last request are kept
// until the arrays have been handed out, so the work is done once.
struct Tables
{
   bool valid = false;
   int key[10];
   int sizes[11];
   std::vector<int> ed, dm, bdr, ltdof, nbrs, shr_off, shr_ltdof, ext_off, ext_ldof;
   std::vector<double> ev;
   std::vector<int64_t> lat;
   void clear() { *this = Tables(); }
};
thread_local Tables g_tables;

} // namespace

static int cartesian_partition_impl(int dim, int nx, int ny, int nz, int p, int sfc_ordering,
                                         int px, int py, int pz, int rank, int *sizes_h,
                                         int *elem_dof_h, int *dof_map_h, int *bdr_ldofs_h,
                                         double *elem_vertices_h, int *ltdof_ldof_h,
                                         int *nbr_rank_h, int *shr_offsets_h, int *shr_ltdof_h,
                                         int *ext_offsets_h, int *ext_ldof_h,
                                         int64_t *ldof_lattice_h)
{
   using b2f::set_error;
   if (dim == 2) { nz = 1; pz = 1; }
   if ((dim != 2 && dim != 3) || !sizes_h || px < 1 || py < 1 || pz < 1 || px > nx || py > ny ||
       pz > nz || p < 1 || rank < 0 || rank >= px * py * pz)
   {
      set_error("b2f_cartesian_partition_h: bad arguments (grid %dx%dx%d over %dx%dx%d elements, "
                "rank %d)", px, py, pz, nx, ny, nz, rank);
      return B2F_ERR_ARG;
   }
   const int key[10] = {dim, nx, ny, nz, p, sfc_ordering, px, py, pz, rank};
   Tables &T = g_tables;
   auto hand_out = [&]() -> int
   {
      std::copy(T.sizes, T.sizes + 11, sizes_h);
      if (elem_dof_h) { std::copy(T.ed.begin(), T.ed.end(), elem_dof_h); }
      if (dof_map_h) { std::copy(T.dm.begin(), T.dm.end(), dof_map_h); }
      if (bdr_ldofs_h) { std::copy(T.bdr.begin(), T.bdr.end(), bdr_ldofs_h); }
      if (elem_vertices_h) { std::copy(T.ev.begin(), T.ev.end(), elem_vertices_h); }
      if (ltdof_ldof_h) { std::copy(T.ltdof.begin(), T.ltdof.end(), ltdof_ldof_h); }
      if (nbr_rank_h) { std::copy(T.nbrs.begin(), T.nbrs.end(), nbr_rank_h); }
      if (ldof_lattice_h) { std::copy(T.lat.begin(), T.lat.end(), ldof_lattice_h); }
      if (shr_offsets_h) { std::copy(T.shr_off.begin(), T.shr_off.end(), shr_offsets_h); }
      if (shr_ltdof_h) { std::copy(T.shr_ltdof.begin(), T.shr_ltdof.end(), shr_ltdof_h); }
      if (ext_offsets_h) { std::copy(T.ext_off.begin(), T.ext_off.end(), ext_offsets_h); }
      if (ext_ldof_h) { std::copy(T.ext_ldof.begin(), T.ext_ldof.end(), ext_ldof_h); }
      if (elem_dof_h) { T.clear(); } // the big arrays have been delivered: release them
      return B2F_OK;
   };
   if (T.valid && std::equal(key, key + 10, T.key)) { return hand_out(); }
   T.clear();
   const bool timing = getenv("B2F_HOST_TIMING") != nullptr;
   auto t_last = std::chrono::steady_clock::now();
   auto lap = [&](const char *what)
   {
      if (!timing) { return; }
      const auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "  partition: %-28s %.2f s\n", what, std::chrono::duration<double>(now - t_last).count());
      t_last = now;
   };
   const Split S[3] = {{nx, px}, {ny, py}, {nz, pz}};
   const int P[3] = {px, py, pz};
   const int b[3] = {rank % px, (rank / px) % py, rank / (px * py)};
   int o[3], ln[3];
   for (int a = 0; a < 3; a++) { o[a] = S[a].lo(b[a]); ln[a] = S[a].lo(b[a] + 1) - o[a]; }

   int ne = 0, nl = 0, nbdr_local = 0;
   int rc = b2f_cartesian_h1_space_h(dim, ln[0], ln[1], ln[2], p, sfc_ordering, &ne, &nl, nullptr,
                                     nullptr, nullptr, nullptr, nullptr);
   if (rc) { return rc; }
   const int p1 = p + 1, nv = dim == 2 ? 4 : 8;
   const int dof = dim == 2 ? p1 * p1 : p1 * p1 * p1;
   std::vector<int> ed((size_t)ne * dof), dm(dof);
   std::vector<double> ev((size_t)ne * nv * dim);
   rc = b2f_cartesian_h1_space_h(dim, ln[0], ln[1], ln[2], p, sfc_ordering, &ne, &nl, ed.data(),
                                 dm.data(), &nbdr_local, nullptr, ev.data());
   if (rc) { return rc; }
   lap("local space numbering");

   // global lattice point of every local dof
   const int64_t L[3] = {(int64_t)p * nx + 1, (int64_t)p * ny + 1, (int64_t)p * nz + 1};
   std::vector<int64_t> lat(nl, -1);
   // (a dof shared by several elements receives the same value from each of them)
   b2f::parallel_for(ne, [&](int64_t eb, int64_t ee)
   {
   for (int e = (int)eb; e < (int)ee; e++)
   {
      int c[3] = {0, 0, 0};
      for (int a = 0; a < dim; a++) { c[a] = (int)std::lround(ev[(size_t)e * nv * dim + nv * a] * ln[a]); }
      for (int k = 0; k < (dim == 3 ? p1 : 1); k++)
         for (int j = 0; j < p1; j++)
            for (int i = 0; i < p1; i++)
            {
               const int lex = i + p1 * (j + p1 * k);
               const int l = ed[(size_t)e * dof + dm[lex]];
               const int64_t g[3] = {(int64_t)p * (o[0] + c[0]) + i, (int64_t)p * (o[1] + c[1]) + j,
                                     (int64_t)p * (o[2] + c[2]) + k};
               __atomic_store_n(&lat[l], g[0] + L[0] * (g[1] + L[1] * g[2]), __ATOMIC_RELAXED);
            }
   }
   });
   lap("lattice points");
   // sharing ranks of a lattice point: per axis, the block(s) whose closed extent holds it
   auto axis_blocks = [&](int a, int64_t g, int out[2]) -> int
   {
      if (a >= dim) { out[0] = 0; return 1; }
      int n = 0;
      // this rank's block holds g by construction; a neighbour holds it only on an interface
      if (b[a] > 0 && g == (int64_t)p * o[a]) { out[n++] = b[a] - 1; }
      out[n++] = b[a];
      if (b[a] + 1 < P[a] && g == (int64_t)p * (o[a] + ln[a])) { out[n++] = b[a] + 1; }
      return n; // n <= 2 unless the block is a single lattice layer thick, which p>=1, ln>=1 excludes
   };
   std::vector<int> owner(nl);
   std::map<int, std::vector<std::pair<int64_t, int>>> shr, ext; // neighbour -> (lattice, ldof)
   std::vector<int> bdr;
   // owner (lowest sharing rank) of every local dof; dofs that are shared or on the global
   // boundary -- a thin surface layer -- are flagged and listed serially afterwards
   std::vector<unsigned char> flag(nl, 0);
   auto classify = [&](int l, int *sharers, int &ns, bool &on_bdr) -> int
   {
      const int64_t g[3] = {lat[l] % L[0], (lat[l] / L[0]) % L[1], lat[l] / (L[0] * L[1])};
      int bl[3][3], nb[3];
      for (int a = 0; a < 3; a++) { nb[a] = axis_blocks(a, g[a], bl[a]); }
      int own = rank;
      ns = 0;
      if (nb[0] * nb[1] * nb[2] > 1) // (interior dofs, the vast majority, have no sharer)
      {
         for (int iz = 0; iz < nb[2]; iz++)
            for (int iy = 0; iy < nb[1]; iy++)
               for (int ix = 0; ix < nb[0]; ix++)
               {
                  const int r = bl[0][ix] + px * (bl[1][iy] + py * bl[2][iz]);
                  if (r != rank) { sharers[ns++] = r; }
                  own = std::min(own, r);
               }
      }
      on_bdr = false;
      for (int a = 0; a < dim; a++) { on_bdr = on_bdr || g[a] == 0 || g[a] == L[a] - 1; }
      return own;
   };
   b2f::parallel_for(nl, [&](int64_t lb, int64_t le)
   {
      int sharers[27], ns;
      bool on_bdr;
      for (int l = (int)lb; l < (int)le; l++)
      {
         owner[l] = classify(l, sharers, ns, on_bdr);
         flag[l] = (unsigned char)((ns > 0) | (on_bdr << 1));
      }
   });
   for (int l = 0; l < nl; l++)
   {
      if (!flag[l]) { continue; }
      int sharers[27], ns;
      bool on_bdr;
      const int own = classify(l, sharers, ns, on_bdr);
      if (own == rank) { for (int k = 0; k < ns; k++) { shr[sharers[k]].push_back({lat[l], l}); } }
      else { ext[own].push_back({lat[l], l}); }
      if (on_bdr) { bdr.push_back(l); }
   }
   lap("owners and neighbour lists");
   // Local renumbering: owned (true) dofs first, in their original relative order, then the
   // dofs owned by neighbours.  With it ltdof_ldof is the identity, a true-dof vector is the
   // head of a local vector, and P / P^T only touch the tail (halo.cu works in place).
   // Element order: interior elements (no dof owned by a neighbour) first, in their original
   // relative order, then the boundary layer -- the operator overlaps the exchange with the
   // former (b2f_operator_set_interior).  Both permutations are applied in one pass.
   int n_interior = 0;
   std::vector<int> l2t(nl, -1), ltdof;
   {
      std::vector<int> perm(nl);
      int nown = 0, next = 0;
      for (int l = 0; l < nl; l++) { nown += owner[l] == rank; }
      next = nown;
      nown = 0;
      for (int l = 0; l < nl; l++) { perm[l] = owner[l] == rank ? nown++ : next++; }
      std::vector<char> is_bdr(ne, 0);
      b2f::parallel_for(ne, [&](int64_t eb, int64_t ee)
      {
         for (int64_t e = eb; e < ee; e++)
         {
            for (int i = 0; i < dof; i++) { if (owner[ed[(size_t)e * dof + i]] != rank) { is_bdr[e] = 1; break; } }
         }
      });
      for (int e = 0; e < ne; e++) { n_interior += !is_bdr[e]; }
      std::vector<int> epos(ne);
      {
         int a = 0, bb = n_interior;
         for (int e = 0; e < ne; e++) { epos[e] = is_bdr[e] ? bb++ : a++; }
      }
      std::vector<int> ed2(ed.size());
      std::vector<double> ev2(ev.size());
      b2f::parallel_for(ne, [&](int64_t eb, int64_t ee)
      {
         for (int64_t e = eb; e < ee; e++)
         {
            const size_t t = (size_t)epos[e];
            for (int i = 0; i < dof; i++) { ed2[t * dof + i] = perm[ed[(size_t)e * dof + i]]; }
            std::copy(ev.begin() + (size_t)e * nv * dim, ev.begin() + (size_t)(e + 1) * nv * dim,
                      ev2.begin() + t * nv * dim);
         }
      });
      ed.swap(ed2);
      ev.swap(ev2);
      std::vector<int64_t> lat2(nl);
      std::vector<int> own2(nl);
      b2f::parallel_for(nl, [&](int64_t lb, int64_t le)
      {
         for (int64_t l = lb; l < le; l++) { lat2[perm[l]] = lat[l]; own2[perm[l]] = owner[l]; }
      });
      lat.swap(lat2);
      owner.swap(own2);
      for (auto &v : bdr) { v = perm[v]; }
      std::sort(bdr.begin(), bdr.end());
      for (auto &kv : shr) { for (auto &q : kv.second) { q.second = perm[q.second]; } }
      for (auto &kv : ext) { for (auto &q : kv.second) { q.second = perm[q.second]; } }
   }
   for (int l = 0; l < nl; l++)
   {
      if (owner[l] == rank) { l2t[l] = (int)ltdof.size(); ltdof.push_back(l); }
   }
   lap("renumbering and element order");
   std::vector<int> nbrs;
   for (auto &kv : shr) { nbrs.push_back(kv.first); }
   for (auto &kv : ext) { nbrs.push_back(kv.first); }
   std::sort(nbrs.begin(), nbrs.end());
   nbrs.erase(std::unique(nbrs.begin(), nbrs.end()), nbrs.end());
   int n_shr = 0, n_ext = 0;
   for (auto &kv : shr) { std::sort(kv.second.begin(), kv.second.end()); n_shr += (int)kv.second.size(); }
   for (auto &kv : ext) { std::sort(kv.second.begin(), kv.second.end()); n_ext += (int)kv.second.size(); }

   T.sizes[0] = ln[0]; T.sizes[1] = ln[1]; T.sizes[2] = ln[2]; T.sizes[3] = ne; T.sizes[4] = nl;
   T.sizes[5] = (int)ltdof.size(); T.sizes[6] = (int)nbrs.size(); T.sizes[7] = n_shr;
   T.sizes[8] = n_ext; T.sizes[9] = (int)bdr.size(); T.sizes[10] = n_interior;
   {
      // global coordinates the way the reference computes them: ((real_t) i / n) * 1.0
      const int N[3] = {nx, ny, nz};
      for (int e = 0; e < ne; e++)
         for (int a = 0; a < dim; a++)
            for (int k = 0; k < nv; k++)
            {
               const size_t at = (size_t)e * nv * dim + nv * a + k;
               const int i = o[a] + (int)std::lround(ev[at] * ln[a]);
               ev[at] = ((double)i / N[a]) * 1.0;
            }
   }
   T.shr_off.assign(nbrs.size() + 1, 0);
   T.ext_off.assign(nbrs.size() + 1, 0);
   T.shr_ltdof.reserve(n_shr);
   T.ext_ldof.reserve(n_ext);
   for (size_t k = 0; k < nbrs.size(); k++)
   {
      T.shr_off[k] = (int)T.shr_ltdof.size();
      T.ext_off[k] = (int)T.ext_ldof.size();
      auto sf = shr.find(nbrs[k]);
      if (sf != shr.end()) { for (auto &q : sf->second) { T.shr_ltdof.push_back(l2t[q.second]); } }
      auto xf = ext.find(nbrs[k]);
      if (xf != ext.end()) { for (auto &q : xf->second) { T.ext_ldof.push_back(q.second); } }
   }
   T.shr_off[nbrs.size()] = (int)T.shr_ltdof.size();
   T.ext_off[nbrs.size()] = (int)T.ext_ldof.size();
   T.ed.swap(ed); T.dm.swap(dm); T.bdr.swap(bdr); T.ev.swap(ev); T.ltdof.swap(ltdof); T.nbrs.swap(nbrs);
   T.lat.swap(lat);
   std::copy(key, key + 10, T.key);
   T.valid = true;
   lap("tables");
   return hand_out();
}

// nothing may throw across the C ABI: multi-GB std::vector growth and std::thread creation can
extern "C" int b2f_cartesian_partition_h(int dim, int nx, int ny, int nz, int p, int sfc_ordering,
                                         int px, int py, int pz, int rank, int *sizes_h,
                                         int *elem_dof_h, int *dof_map_h, int *bdr_ldofs_h,
                                         double *elem_vertices_h, int *ltdof_ldof_h,
                                         int *nbr_rank_h, int *shr_offsets_h, int *shr_ltdof_h,
                                         int *ext_offsets_h, int *ext_ldof_h,
                                         int64_t *ldof_lattice_h)
{
   try
   {
      return cartesian_partition_impl(dim, nx, ny, nz, p, sfc_ordering, px, py, pz, rank, sizes_h, elem_dof_h,
                                      dof_map_h, bdr_ldofs_h, elem_vertices_h, ltdof_ldof_h, nbr_rank_h,
                                      shr_offsets_h, shr_ltdof_h, ext_offsets_h, ext_ldof_h, ldof_lattice_h);
   }
   catch (const std::exception &ex)
   {
      b2f::set_error("b2f_cartesian_partition_h: %s", ex.what());
      return B2F_ERR_ARG;
   }
}
