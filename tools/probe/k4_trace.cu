// Where a K4 tile's lifetime goes: per-tile globaltimer stamps (loop top, ticket, counts + scan, look-back done, scatter
// done), look-back rounds and spins.  Builds the library's own kernel with -DSEMGATE_K4_TRACE.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DSEMGATE_K4_TRACE -DSEMGATE_BUILD \
//        -I multi-level-indoor-slam_b200/csrc -o tools/probe/_k4_trace tools/probe/k4_trace.cu
//   tools/probe/_k4_trace [rows] [k] [entries per row]
#include "../../multi-level-indoor-slam_b200/csrc/kernels.cu"

#include <cstdio>
#include <vector>
#include <algorithm>

int main(int argc, char** argv) {
  const int64_t Q = argc > 1 ? atoll(argv[1]) : 1000000;
  const int k = argc > 2 ? atoi(argv[2]) : 25;
  const int fill = argc > 3 ? atoi(argv[3]) : 25;
  float* sc; int32_t* ix; uint8_t* va; int32_t* ct;
  int32_t *oq, *om; float* os; uint8_t* ov; int64_t* tot; void* ws;
  cudaMalloc(&sc, Q * k * 4); cudaMalloc(&ix, Q * k * 4); cudaMalloc(&va, Q * k); cudaMalloc(&ct, Q * 4);
  cudaMalloc(&oq, Q * k * 4); cudaMalloc(&om, Q * k * 4); cudaMalloc(&os, Q * k * 4); cudaMalloc(&ov, Q * k); cudaMalloc(&tot, 8);
  cudaMalloc(&ws, semgate::compact_workspace_bytes(Q));
  cudaMemset(sc, 0, Q * k * 4); cudaMemset(ix, 0, Q * k * 4); cudaMemset(va, 1, Q * k);
  std::vector<int32_t> h(Q, fill);
  cudaMemcpy(ct, h.data(), Q * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    int rc = semgate::launch_compact(sc, ix, va, ct, Q, k, false, 0, oq, om, os, ov, tot, ws, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e != cudaSuccess) { printf("launch failed %d %s\n", rc, cudaGetErrorString(e)); return 1; }
    cudaEventElapsedTime(&ms, e0, e1);
  }
  const int64_t tiles = std::min<int64_t>((Q + 127) / 128, 65536);
  std::vector<unsigned long long> tr(8 * tiles);
  cudaMemcpyFromSymbol(tr.data(), semgate::k4_trace, tr.size() * 8);
  double s[4] = {0, 0, 0, 0}, rounds = 0, spins = 0;
  unsigned long long tmin = ~0ull, tmax = 0, maxr = 0;
  for (int64_t t = 1; t < tiles; ++t) {
    const unsigned long long* p = &tr[8 * t];
    s[0] += double(p[1] - p[0]); s[1] += double(p[2] - p[1]); s[2] += double(p[3] - p[2]); s[3] += double(p[4] - p[3]);
    rounds += double(p[5]); spins += double(p[6]); maxr = std::max(maxr, p[5]);
    tmin = std::min(tmin, p[0]); tmax = std::max(tmax, p[4]);
  }
  const double n = double(tiles - 1);
  printf("rows %lld k %d entries/row %d: kernel+memset %.1f us (events), span of the tiles' stamps %.1f us\n", (long long)Q, k, fill, ms * 1e3, (tmax - tmin) * 1e-3);
  printf("per tile (us): ticket %.2f | counts + scan %.2f | look-back %.2f | scatter %.2f | look-back rounds %.1f (max %llu), spins %.1f\n",
         s[0] / n * 1e-3, s[1] / n * 1e-3, s[2] / n * 1e-3, s[3] / n * 1e-3, rounds / n, maxr, spins / n);
  return 0;
}
