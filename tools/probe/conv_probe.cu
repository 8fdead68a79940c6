// Probe (kept as evidence for DESIGN.md 6): do full-mask votes inside a persistent loop see all 32 lanes after divergent work? (yes: 0 partial ballots on B200)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __noinline__ float slow_div(float a, float b) { return a / b; }

template <int kVariant>
__global__ void probe(unsigned* out, int trips, float* sink) {
  unsigned lane = threadIdx.x & 31;
  unsigned rng = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  int state = 0;
  float acc = 1.0f + lane;
  unsigned bad_ballot = 0, bad_active = 0;
  for (int t = 0; t < trips; ++t) {
    if (kVariant == 2) __syncwarp();
    unsigned b = __ballot_sync(0xffffffffu, true);
    if (b != 0xffffffffu) bad_ballot++;
    if (__activemask() != 0xffffffffu) bad_active++;
    // divergent section A: switch with different amounts of work, including IEEE division
#pragma unroll 1
    for (int round = 0; round < 4; ++round) {
      const bool slow = state >= 1 && state <= 4;
      if (!__any_sync(0xffffffffu, slow)) break;
      if (!slow) continue;
      switch (state) {
        case 1: for (int i = 0; i < (int)(rng & 63); ++i) acc = slow_div(acc + 1.0f, 1.0001f + i); state = 2; break;
        case 2: acc = sqrtf(acc + 2.0f); state = (rng & 4) ? 3 : 5; break;
        case 3: for (int i = 0; i < (int)((rng >> 8) & 15); ++i) acc = expf(-acc) + 1.0f; state = 4; break;
        case 4: acc = acc / (acc + 3.0f); state = 5; break;
        default: break;
      }
      rng = rng * 1664525u + 1013904223u;
    }
    // divergent section B: "march" loop with vote-controlled exit and continue
    for (int it = 0; it < 32; ++it) {
      const bool m = state == 5;
      const int marchers = __popc(__ballot_sync(0xffffffffu, m));
      if (marchers == 0 || (it >= 4 && marchers < 12)) break;
      if (!m) continue;
      rng = rng * 1664525u + 1013904223u;
      int n = (rng >> 10) & 15;
      for (int j = 0; j < n; ++j) acc += 0.001f;
      if ((rng & 0x70000) == 0) { state = 0; continue; }
    }
    if (state == 0) { rng = rng * 1664525u + 1013904223u; state = 1 + (rng >> 30); }
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  atomicAdd(&out[0], bad_ballot);
  atomicAdd(&out[1], bad_active);
}

int main() {
  unsigned* d; float* sink;
  cudaMalloc(&d, 8); cudaMalloc(&sink, 296 * 128 * 4);
  for (int v = 0; v < 3; ++v) {
    cudaMemset(d, 0, 8);
    if (v == 0) probe<0><<<296, 128>>>(d, 20000, sink);
    if (v == 1) probe<1><<<296, 128>>>(d, 20000, sink);
    if (v == 2) probe<2><<<296, 128>>>(d, 20000, sink);
    unsigned h[2];
    cudaError_t e = cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("variant %d: err=%s partial ballots=%u partial activemask=%u\n", v, cudaGetErrorString(e), h[0], h[1]);
  }
  return 0;
}
