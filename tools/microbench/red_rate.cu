// red_rate.cu -- how many FP64 global reductions (RED.E.ADD.F64) per second does the chip retire?
// The fused operator scatters D^3 values per element with them; this measures the ceiling for
// (a) an H1-like pattern (every address hit ~ (D/p)^3 times, element-local runs), (b) private
// contiguous addresses, (c) plain stores of the same data as the no-atomic reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/red_rate tools/microbench/red_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

template <int MODE>
__global__ void scatter(const int *__restrict__ idx, double *y, long long n)
{
   for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
   {
      const int g = idx[i];
      if (MODE == 0) { atomicAdd(y + g, 1.0); }
      else { y[g] = 1.0; }
   }
}

int main(int argc, char **argv)
{
   const int p = argc > 1 ? atoi(argv[1]) : 4, n = argc > 2 ? atoi(argv[2]) : 64;
   const int D = p + 1, nd1 = n * p + 1;
   const long long ne = 1LL * n * n * n, dof = 1LL * D * D * D, N = ne * dof, nd = 1LL * nd1 * nd1 * nd1;
   std::vector<int> h1((size_t)N), dg((size_t)N);
   for (long long e = 0; e < ne; e++)
   {
      const int ex = e % n, ey = (e / n) % n, ez = e / (1LL * n * n);
      for (int k = 0; k < dof; k++)
      {
         const int dx = k % D, dy = (k / D) % D, dz = k / (D * D);
         h1[e * dof + k] = (int)((ex * p + dx) + 1LL * nd1 * ((ey * p + dy) + 1LL * nd1 * (ez * p + dz)));
         dg[e * dof + k] = (int)(e * dof + k);
      }
   }
   int *d_idx; double *y;
   cudaMalloc(&d_idx, sizeof(int) * N);
   cudaMalloc(&y, sizeof(double) * N);
   cudaMemset(y, 0, sizeof(double) * N);
   cudaEvent_t a, b;
   cudaEventCreate(&a); cudaEventCreate(&b);
   const char *names[2] = {"lexicographic H1 lattice", "private contiguous (dg)"};
   for (int pat = 0; pat < 2; pat++)
   {
      cudaMemcpy(d_idx, pat ? dg.data() : h1.data(), sizeof(int) * N, cudaMemcpyHostToDevice);
      for (int mode = 0; mode < 2; mode++)
      {
         float best = 1e30f;
         for (int r = 0; r < 6; r++)
         {
            cudaEventRecord(a);
            if (mode == 0) { scatter<0><<<148 * 16, 256>>>(d_idx, y, N); }
            else { scatter<1><<<148 * 16, 256>>>(d_idx, y, N); }
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (r && ms < best) { best = ms; }
         }
         printf("p=%d n=%d %-26s %-6s: %8.4f ms  %7.1f G/s  (%lld scattered values, %lld distinct)\n", p, n,
                names[pat], mode ? "store" : "RED", best, N / best / 1e6, N, pat ? N : nd);
      }
   }
   printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
   return 0;
}
