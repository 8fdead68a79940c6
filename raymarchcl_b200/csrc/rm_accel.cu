// rm_accel.cu -- occupancy acceleration data derived from the uploaded uint8 volume for one
// isoVal. Nothing here changes results: the fast kernel uses these tables only to decide which
// voxel fetches of the reference's fixed-step march (renderer.cl:219-234) can be elided because
// their outcome (not solid) is already known.
//
//   solid  bit-bricks of (v >  isoVal), the march predicate (renderer.cl:222):   one 64-bit word
//   occ    bit-bricks of (v >= isoVal), the normal predicate (renderer.cl:176):  per 4x4x4 voxels,
//          bit = (x&3) | (y&3)<<2 | (z&3)<<4, words laid out x-fastest over the brick grid
//   dist   one byte per macro-cell (cell = 1<<cell_shift voxels, >= one brick): Chebyshev distance
//          in cells to the nearest cell that contains a solid voxel, saturated at RM_DIST_CAP.
//          All cells within Chebyshev distance dist-1 are free of solid voxels.
#include "rm_kernels.h"

namespace {

// one thread per brick; adjacent threads read adjacent 4-byte runs of a voxel row (coalesced)
__global__ void __launch_bounds__(256)
k_build_bricks(const uint8_t* __restrict__ vox, int rx, int ry, int rz, int iso, int bx, int by, int bz,
               uint64_t* __restrict__ solid, uint64_t* __restrict__ occ, unsigned* __restrict__ differ) {
  const long long b = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long nb = (long long)bx * by * bz;
  if (b >= nb) return;
  const int ix = (int)(b % bx);
  const int iy = (int)((b / bx) % by);
  const int iz = (int)(b / ((long long)bx * by));
  const int x0 = ix * 4, y0 = iy * 4, z0 = iz * 4;
  const bool fast_row = (rx & 3) == 0 && ((uintptr_t)vox & 3) == 0;
  uint64_t ws = 0, wo = 0;
#pragma unroll
  for (int dz = 0; dz < 4; ++dz) {
    const int z = z0 + dz;
    if (z >= rz) break;
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
      const int y = y0 + dy;
      if (y >= ry) break;
      const size_t row = ((size_t)z * ry + y) * rx + x0;
      uint32_t v4 = 0;
      if (fast_row) {
        v4 = __ldg(reinterpret_cast<const uint32_t*>(vox + row));
      } else {
        for (int dx = 0; dx < 4; ++dx)
          if (x0 + dx < rx) v4 |= (uint32_t)__ldg(vox + row + dx) << (8 * dx);
          else v4 |= 0u;
      }
      uint32_t s4 = 0, o4 = 0;
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const int v = (v4 >> (8 * dx)) & 255;
        const bool in = x0 + dx < rx;
        s4 |= (uint32_t)(in && v > iso) << dx;
        o4 |= (uint32_t)(in && v >= iso) << dx;
      }
      const int sh = dy * 4 + dz * 16;
      ws |= (uint64_t)s4 << sh;
      wo |= (uint64_t)o4 << sh;
    }
  }
  solid[b] = ws;
  occ[b] = wo;
  if (ws != wo) atomicOr(differ, 1u);
}

// cell occupancy: 0 when any brick of the macro-cell holds a solid voxel, else RM_DIST_CAP
__global__ void __launch_bounds__(256)
k_cell_seed(const uint64_t* __restrict__ solid, int bx, int by, int bz, int mx, int my, int mz,
            int bricks_per_cell, uint8_t* __restrict__ out) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  if (c >= (long long)mx * my * mz) return;
  const int cx = (int)(c % mx), cy = (int)((c / mx) % my), cz = (int)(c / ((long long)mx * my));
  bool any = false;
  for (int dz = 0; dz < bricks_per_cell && !any; ++dz) {
    const int z = cz * bricks_per_cell + dz;
    if (z >= bz) break;
    for (int dy = 0; dy < bricks_per_cell && !any; ++dy) {
      const int y = cy * bricks_per_cell + dy;
      if (y >= by) break;
      for (int dx = 0; dx < bricks_per_cell; ++dx) {
        const int x = cx * bricks_per_cell + dx;
        if (x >= bx) break;
        if (__ldg(solid + ((size_t)z * by + y) * bx + x) != 0) { any = true; break; }
      }
    }
  }
  out[c] = any ? 0 : RM_DIST_CAP;
}

// One axis of the separable Chebyshev transform:
//   out(p) = min over k in [-CAP, CAP] of max(|k|, in(p + k*axis)),   clamped to CAP.
// Applied along x, then y, then z to the 0/CAP seed this yields min_q max(|dx|,|dy|,|dz|).
__global__ void __launch_bounds__(256)
k_cheb_axis(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int mx, int my, int mz, int axis) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  if (c >= (long long)mx * my * mz) return;
  const int cx = (int)(c % mx), cy = (int)((c / mx) % my), cz = (int)(c / ((long long)mx * my));
  const int pos = axis == 0 ? cx : (axis == 1 ? cy : cz);
  const int len = axis == 0 ? mx : (axis == 1 ? my : mz);
  const long long stride = axis == 0 ? 1 : (axis == 1 ? mx : (long long)mx * my);
  int best = __ldg(in + c);
  for (int k = 1; k < best; ++k) {  // a candidate at offset k can only give max(k, .) >= k
    int a = RM_DIST_CAP, b = RM_DIST_CAP;
    if (pos - k >= 0) a = __ldg(in + c - k * stride);
    if (pos + k < len) b = __ldg(in + c + k * stride);
    const int m = max(k, min(a, b));
    best = min(best, m);
  }
  out[c] = (uint8_t)best;
}

// 4 bits per cell: min(dist, 15), two cells per byte (low nibble = even cell)
__global__ void __launch_bounds__(256)
k_pack_nibbles(const uint8_t* __restrict__ dist, long long cells, uint8_t* __restrict__ nib, unsigned nib_bytes) {
  const long long b = (long long)blockIdx.x * 256 + threadIdx.x;
  if (b >= nib_bytes) return;
  const long long c = 2 * b;
  const int lo = c < cells ? min((int)__ldg(dist + c), 15) : 0;
  const int hi = c + 1 < cells ? min((int)__ldg(dist + c + 1), 15) : 0;
  nib[b] = (uint8_t)(lo | (hi << 4));
}

}  // namespace

cudaError_t rm_accel_build(const uint8_t* d_vox, int rx, int ry, int rz, int iso, int cell_shift,
                           RmAccelStorage* st, cudaStream_t stream) {
  RmAccel& a = st->view;
  a.vox = d_vox;
  a.bx = (rx + 3) >> 2; a.by = (ry + 3) >> 2; a.bz = (rz + 3) >> 2;
  a.cell_shift = cell_shift < 2 ? 2 : cell_shift;
  const int cell = 1 << a.cell_shift;
  a.cellf = (float)cell;
  a.rxf = (float)rx; a.ryf = (float)ry; a.rzf = (float)rz;
  a.inv_rxf = 1.0f / a.rxf; a.inv_ryf = 1.0f / a.ryf; a.inv_rzf = 1.0f / a.rzf;
  a.pow2 = ((rx & (rx - 1)) == 0 && (ry & (ry - 1)) == 0 && (rz & (rz - 1)) == 0) ? 1 : 0;
  a.mx = (rx + cell - 1) >> a.cell_shift; a.my = (ry + cell - 1) >> a.cell_shift; a.mz = (rz + cell - 1) >> a.cell_shift;
  const size_t nb = (size_t)a.bx * a.by * a.bz, nc = (size_t)a.mx * a.my * a.mz;
  cudaError_t e;
  if (st->brick_capacity < nb) {
    cudaFree(st->d_solid); cudaFree(st->d_occ);
    st->d_solid = st->d_occ = nullptr; st->brick_capacity = 0;
    if ((e = cudaMalloc(&st->d_solid, nb * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&st->d_occ, nb * 8)) != cudaSuccess) return e;
    st->brick_capacity = nb;
  }
  if (st->cell_capacity < nc) {
    cudaFree(st->d_dist); cudaFree(st->d_tmp); cudaFree(st->d_nib);
    st->d_dist = st->d_tmp = st->d_nib = nullptr; st->cell_capacity = 0;
    if ((e = cudaMalloc(&st->d_dist, nc)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&st->d_tmp, nc)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&st->d_nib, (nc / 2 + 16) & ~(size_t)15)) != cudaSuccess) return e;
    st->cell_capacity = nc;
  }
  if (!st->d_flag && (e = cudaMalloc(&st->d_flag, sizeof(unsigned))) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(st->d_flag, 0, sizeof(unsigned), stream)) != cudaSuccess) return e;
  k_build_bricks<<<(unsigned)((nb + 255) / 256), 256, 0, stream>>>(d_vox, rx, ry, rz, iso, a.bx, a.by, a.bz,
                                                                  st->d_solid, st->d_occ, st->d_flag);
  const unsigned cb = (unsigned)((nc + 255) / 256);
  k_cell_seed<<<cb, 256, 0, stream>>>(st->d_solid, a.bx, a.by, a.bz, a.mx, a.my, a.mz, cell >> 2, st->d_dist);
  k_cheb_axis<<<cb, 256, 0, stream>>>(st->d_dist, st->d_tmp, a.mx, a.my, a.mz, 0);
  k_cheb_axis<<<cb, 256, 0, stream>>>(st->d_tmp, st->d_dist, a.mx, a.my, a.mz, 1);
  k_cheb_axis<<<cb, 256, 0, stream>>>(st->d_dist, st->d_tmp, a.mx, a.my, a.mz, 2);
  const unsigned nib_bytes = (unsigned)(((nc + 1) / 2 + 15) & ~(size_t)15);
  k_pack_nibbles<<<(nib_bytes + 255) / 256, 256, 0, stream>>>(st->d_tmp, (long long)nc, st->d_nib, nib_bytes);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  // (round 1 read the `differ` flag back here to alias occ to solid when no voxel equals isoVal; that
  //  cost a stream synchronisation on every upload -- the build is now fully asynchronous and the normal
  //  taps, a few dozen per pixel-sample, always read the occ bricks)
  a.solid = st->d_solid;
  a.occ = st->d_occ;
  a.dist = st->d_tmp;
  a.nib = st->d_nib;
  a.nib_bytes = nib_bytes;
  st->iso = iso;
  st->valid = true;
  st->launches = 6;
  return cudaSuccess;
}

void rm_accel_free(RmAccelStorage* st) {
  cudaFree(st->d_solid); cudaFree(st->d_occ); cudaFree(st->d_dist); cudaFree(st->d_tmp); cudaFree(st->d_nib); cudaFree(st->d_flag);
  *st = RmAccelStorage{};
}
