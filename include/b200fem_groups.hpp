// include/b200fem_groups.hpp -- from MFEM's group tables to b2f_halo_create's arguments.
//
// Header-only, no MFEM and no MPI headers: plain integer code an MPI build of the adapter calls
// with what ParFiniteElementSpace already holds.  It restates, on plain arrays,
//   * GroupCommunicator::Finalize (general/communication.cpp:368-470): which groups go to / come
//     from which neighbour -- a group I am master of is sent to every other member, a group
//     mastered by rank r is received from r -- and in WHICH ORDER: the sender lists its groups by
//     its own group number, the receiver by the number the group has on the master
//     (GroupTopology::GetGroupMasterGroup), so both sides enumerate a pair's groups identically
//     without communicating;
//   * GroupCommunicator::GetNeighborLTDofTable / GetNeighborLDofTable (communication.cpp:543-612):
//     per neighbour, the concatenated true-dof (send) and local-dof (receive) lists;
//   * DeviceConformingProlongationOperator's constructor (fem/pfespace.cpp:5625-5690): ltdof_ldof
//     from the restriction matrix R (one entry per row).
//
// The MPI adapter then is ("ParFiniteElementSpace binding"):
//     b200::GroupTables g = ...fill from pfes.GroupComm(), pfes.GetGroupTopo(), pfes.GetLocalTDofNumber...
//     b200::HaloTables h = b200::halo_tables_from_groups(g);
//     b2f_halo_create(ctx, id, rank, nranks, g.n_ldofs, h.ltdof_ldof.size(), h.ltdof_ldof.data(),
//                     h.nbr_rank.size(), h.nbr_rank.data(), h.shr_offsets.data(), h.shr_ltdof.data(),
//                     h.ext_offsets.data(), h.ext_ldof.data(), &halo);
#ifndef B200FEM_GROUPS_HPP
#define B200FEM_GROUPS_HPP

#include <algorithm>
#include <stdexcept>
#include <utility>
#include <vector>

namespace b200
{

/// What GroupTopology (general/communication.hpp:145-281) and GroupCommunicator (:283-446) hold for
/// one rank.  Group 0 is the local group (dofs nobody shares); it is ignored.
struct GroupTables
{
   int my_rank = 0;
   int n_groups = 0;
   std::vector<int> group_master_rank;  ///< [n_groups] rank of gtopo.GetGroupMaster(g)
   std::vector<int> group_master_group; ///< [n_groups] gtopo.GetGroupMasterGroup(g): g's number on its master
   std::vector<int> group_rank_offsets; ///< [n_groups+1] into group_ranks
   std::vector<int> group_ranks;        ///< member ranks of each group, my_rank included (gtopo.GetGroup)
   std::vector<int> group_ldof_offsets; ///< [n_groups+1] GroupCommunicator::GroupLDofTable().GetI()
   std::vector<int> group_ldofs;        ///< ... GetJ(): the group's local dofs in the communicator's order
   int n_ldofs = 0;
   std::vector<int> ldof_ltdof;         ///< [n_ldofs] pfes.GetLocalTDofNumber(ldof), -1 if not owned
};

/// b2f_halo_create's table arguments (include/b200fem.h).
struct HaloTables
{
   std::vector<int> ltdof_ldof;  ///< [n_tdofs] owned local dofs, ascending true-dof number
   std::vector<int> nbr_rank;    ///< neighbour ranks, ascending
   std::vector<int> shr_offsets; ///< [n_nbr+1]
   std::vector<int> shr_ltdof;   ///< owned true dofs sent to each neighbour
   std::vector<int> ext_offsets; ///< [n_nbr+1]
   std::vector<int> ext_ldof;    ///< local dofs owned by each neighbour
};

inline HaloTables halo_tables_from_groups(const GroupTables &g)
{
   auto need = [](bool ok, const char *what) { if (!ok) { throw std::invalid_argument(what); } };
   need(g.n_groups >= 0 && (int)g.group_master_rank.size() == g.n_groups &&
        (int)g.group_master_group.size() == g.n_groups, "group master tables: wrong size");
   need((int)g.group_rank_offsets.size() == g.n_groups + 1 && (int)g.group_ldof_offsets.size() == g.n_groups + 1,
        "group offset tables: wrong size");
   need((int)g.ldof_ltdof.size() == g.n_ldofs, "ldof_ltdof: wrong size");
   HaloTables h;
   // ltdof_ldof: the inverse of ldof_ltdof on the owned dofs
   int n_tdofs = 0;
   for (int l = 0; l < g.n_ldofs; l++) { n_tdofs = std::max(n_tdofs, g.ldof_ltdof[l] + 1); }
   h.ltdof_ldof.assign(n_tdofs, -1);
   for (int l = 0; l < g.n_ldofs; l++)
   {
      const int t = g.ldof_ltdof[l];
      if (t >= 0)
      {
         need(h.ltdof_ldof[t] == -1, "ldof_ltdof: two local dofs map to one true dof");
         h.ltdof_ldof[t] = l;
      }
   }
   for (int t = 0; t < n_tdofs; t++) { need(h.ltdof_ldof[t] >= 0, "ldof_ltdof: a true dof has no local dof"); }

   // (neighbour rank, ordering key, group): send lists keyed by my group number, receive lists
   // by the group's number on its master
   std::vector<std::pair<std::pair<int, int>, int>> send, recv;
   for (int gr = 1; gr < g.n_groups; gr++)
   {
      if (g.group_ldof_offsets[gr + 1] == g.group_ldof_offsets[gr]) { continue; } // no dofs in this group
      if (g.group_master_rank[gr] != g.my_rank)
      {
         recv.push_back({{g.group_master_rank[gr], g.group_master_group[gr]}, gr});
         continue;
      }
      for (int k = g.group_rank_offsets[gr]; k < g.group_rank_offsets[gr + 1]; k++)
      {
         const int r = g.group_ranks[k];
         if (r != g.my_rank) { send.push_back({{r, gr}, gr}); }
      }
   }
   std::sort(send.begin(), send.end());
   std::sort(recv.begin(), recv.end());
   for (const auto &s : send) { h.nbr_rank.push_back(s.first.first); }
   for (const auto &r : recv) { h.nbr_rank.push_back(r.first.first); }
   std::sort(h.nbr_rank.begin(), h.nbr_rank.end());
   h.nbr_rank.erase(std::unique(h.nbr_rank.begin(), h.nbr_rank.end()), h.nbr_rank.end());

   const int nn = (int)h.nbr_rank.size();
   h.shr_offsets.assign(nn + 1, 0);
   h.ext_offsets.assign(nn + 1, 0);
   size_t is = 0, ir = 0;
   for (int n = 0; n < nn; n++)
   {
      const int rank = h.nbr_rank[n];
      for (; is < send.size() && send[is].first.first == rank; is++)
      {
         const int gr = send[is].second;
         for (int k = g.group_ldof_offsets[gr]; k < g.group_ldof_offsets[gr + 1]; k++)
         {
            const int t = g.ldof_ltdof[g.group_ldofs[k]];
            need(t >= 0, "a dof of a group I am master of is not a true dof");
            h.shr_ltdof.push_back(t);
         }
      }
      h.shr_offsets[n + 1] = (int)h.shr_ltdof.size();
      for (; ir < recv.size() && recv[ir].first.first == rank; ir++)
      {
         const int gr = recv[ir].second;
         for (int k = g.group_ldof_offsets[gr]; k < g.group_ldof_offsets[gr + 1]; k++)
         {
            const int l = g.group_ldofs[k];
            need(l >= 0 && l < g.n_ldofs && g.ldof_ltdof[l] < 0, "a dof received from its master is marked as owned");
            h.ext_ldof.push_back(l);
         }
      }
      h.ext_offsets[n + 1] = (int)h.ext_ldof.size();
   }
   return h;
}

} // namespace b200
#endif
