// mfem_b200/csrc/host_parallel.h -- plain std::thread fan-out for the few host loops over all
// E-dofs / L-dofs (hundreds of millions of independent iterations at benchmark sizes); no OpenMP
// runtime is pulled into the host process.  B2F_HOST_THREADS overrides the thread count (<= 16),
// B2F_HOST_PAR_MIN the loop length below which a loop stays serial (default 65536; tests set it to
// 1 so that small meshes exercise the threaded paths).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

namespace b2f
{
template <typename F>
void parallel_for(int64_t n, F &&body)
{
   unsigned nt = std::thread::hardware_concurrency();
   if (const char *s = getenv("B2F_HOST_THREADS")) { nt = (unsigned)atoi(s); }
   if (nt > 16) { nt = 16; }
   int64_t nmin = 1 << 16;
   if (const char *s = getenv("B2F_HOST_PAR_MIN")) { nmin = atoll(s); }
   if (nt < 2 || n < nmin) { body((int64_t)0, n); return; }
   std::vector<std::thread> th;
   const int64_t chunk = (n + nt - 1) / nt;
   for (unsigned k = 0; k < nt; k++)
   {
      const int64_t b = (int64_t)k * chunk, e = std::min<int64_t>(n, b + chunk);
      if (b >= e) { break; }
      th.emplace_back([&body, b, e]() { body(b, e); });
   }
   for (auto &t : th) { t.join(); }
}
} // namespace b2f
