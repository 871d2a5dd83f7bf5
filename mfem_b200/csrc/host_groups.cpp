// mfem_b200/csrc/host_groups.cpp -- C entry point over include/b200fem_groups.hpp: the
// ParFiniteElementSpace binding (GroupTopology / GroupCommunicator tables -> b2f_halo_create's
// arguments; general/communication.cpp:368-612, fem/pfespace.cpp:5625-5690) as plain C so that it
// can be exercised without MPI (tests/test_host_cpu.py checks it against the Cartesian
// partitioner's own tables).
#include <algorithm>
#include <exception>

#include "../../include/b200fem.h"
#include "../../include/b200fem_groups.hpp"

namespace b2f
{
void set_error(const char *fmt, ...);
}

extern "C" int b2f_halo_tables_from_groups_h(int my_rank, int n_groups, const int *group_master_rank_h,
                                             const int *group_master_group_h, const int *group_rank_offsets_h,
                                             const int *group_ranks_h, const int *group_ldof_offsets_h,
                                             const int *group_ldofs_h, int n_ldofs, const int *ldof_ltdof_h,
                                             int *sizes_h, int *ltdof_ldof_h, int *nbr_rank_h, int *shr_offsets_h,
                                             int *shr_ltdof_h, int *ext_offsets_h, int *ext_ldof_h)
{
   if (n_groups < 0 || n_ldofs < 0 || !sizes_h || (n_groups > 0 && (!group_master_rank_h || !group_master_group_h ||
       !group_rank_offsets_h || !group_ldof_offsets_h)) || (n_ldofs > 0 && !ldof_ltdof_h))
   {
      b2f::set_error("b2f_halo_tables_from_groups_h: bad arguments");
      return B2F_ERR_ARG;
   }
   try
   {
      b200::GroupTables g;
      g.my_rank = my_rank;
      g.n_groups = n_groups;
      g.n_ldofs = n_ldofs;
      g.group_master_rank.assign(group_master_rank_h, group_master_rank_h + n_groups);
      g.group_master_group.assign(group_master_group_h, group_master_group_h + n_groups);
      if (n_groups > 0)
      {
         g.group_rank_offsets.assign(group_rank_offsets_h, group_rank_offsets_h + n_groups + 1);
         g.group_ldof_offsets.assign(group_ldof_offsets_h, group_ldof_offsets_h + n_groups + 1);
         g.group_ranks.assign(group_ranks_h, group_ranks_h + g.group_rank_offsets[n_groups]);
         g.group_ldofs.assign(group_ldofs_h, group_ldofs_h + g.group_ldof_offsets[n_groups]);
      }
      else
      {
         g.group_rank_offsets.assign(1, 0);
         g.group_ldof_offsets.assign(1, 0);
      }
      g.ldof_ltdof.assign(ldof_ltdof_h, ldof_ltdof_h + n_ldofs);
      const b200::HaloTables h = b200::halo_tables_from_groups(g);
      sizes_h[0] = (int)h.ltdof_ldof.size();
      sizes_h[1] = (int)h.nbr_rank.size();
      sizes_h[2] = (int)h.shr_ltdof.size();
      sizes_h[3] = (int)h.ext_ldof.size();
      if (ltdof_ldof_h) { std::copy(h.ltdof_ldof.begin(), h.ltdof_ldof.end(), ltdof_ldof_h); }
      if (nbr_rank_h) { std::copy(h.nbr_rank.begin(), h.nbr_rank.end(), nbr_rank_h); }
      if (shr_offsets_h) { std::copy(h.shr_offsets.begin(), h.shr_offsets.end(), shr_offsets_h); }
      if (shr_ltdof_h) { std::copy(h.shr_ltdof.begin(), h.shr_ltdof.end(), shr_ltdof_h); }
      if (ext_offsets_h) { std::copy(h.ext_offsets.begin(), h.ext_offsets.end(), ext_offsets_h); }
      if (ext_ldof_h) { std::copy(h.ext_ldof.begin(), h.ext_ldof.end(), ext_ldof_h); }
   }
   catch (const std::exception &ex)
   {
      b2f::set_error("b2f_halo_tables_from_groups_h: %s", ex.what());
      return B2F_ERR_ARG;
   }
   return B2F_OK;
}
