// mfem_b200/csrc/host_parallel.h -- plain std::thread fan-out for the few host loops over all
// E-dofs / L-dofs (hundreds of millions of independent iterations at benchmark sizes); no OpenMP
// runtime is pulled into the host process.  B2F_HOST_THREADS overrides the thread count (<= 16).
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
   if (nt < 2 || n < (1 << 16)) { body((int64_t)0, n); return; }
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
