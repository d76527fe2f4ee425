// Voxel-grid centroid down-sampling of a point cloud: what the reference asks of
// open3d's PointCloud.voxel_down_sample at SDD:2486-2500, 2640-2680 (2 mm scene memory, 25 mm saved
// clouds).  Open3D 0.17 semantics, restated in oracle/geometry_ref.c: the grid is anchored at
// min_bound - voxel/2, a point belongs to voxel floor((p - origin) / voxel) per axis (float64), the
// output is the mean of each voxel's points; the output order is unspecified there.
//
// One pass over the points: a hash table keyed by the packed voxel index (3 x 21 bits) is filled with
// atomicCAS, and every point adds (p - origin) as 2^-36 m fixed point to its voxel's three 64-bit sums,
// so the result does not depend on the order in which threads arrive (float64 atomics would).  The
// centroid differs from a sequential float64 mean by < 2e-11 m.  HBM: 24 B/point read; the table
// (36 B/slot, 2 to 4 slots per point) lives in L2 for the clouds of this pipeline (<= 1 M points).
#include "common.cuh"

namespace prg {

constexpr unsigned long long kVoxEmpty = ~0ull;
constexpr double kVoxFix = 68719476736.0;       // 2^36
constexpr double kVoxMaxIndex = 2097152.0;      // 2^21 voxels per axis

// monotone map double -> uint64 so that atomicMin orders like the doubles do
__device__ __forceinline__ unsigned long long vox_ordered(double d) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double vox_unordered(unsigned long long o) {
  const unsigned long long u = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
  return __longlong_as_double((long long)u);
}
__device__ __forceinline__ unsigned long long vox_mix(unsigned long long x) {   // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

__global__ void __launch_bounds__(256)
k_vox_bounds(const double* __restrict__ pts, long long n, unsigned long long* __restrict__ minb) {
  unsigned long long m0 = kVoxEmpty, m1 = kVoxEmpty, m2 = kVoxEmpty;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long a = vox_ordered(pts[i * 3 + 0]);
    const unsigned long long b = vox_ordered(pts[i * 3 + 1]);
    const unsigned long long c = vox_ordered(pts[i * 3 + 2]);
    m0 = a < m0 ? a : m0;
    m1 = b < m1 ? b : m1;
    m2 = c < m2 ? c : m2;
  }
  if (m0 != kVoxEmpty) {
    atomicMin(minb + 0, m0);
    atomicMin(minb + 1, m1);
    atomicMin(minb + 2, m2);
  }
}

__global__ void __launch_bounds__(256)
k_vox_insert(const double* __restrict__ pts, long long n, double voxel,
             const unsigned long long* __restrict__ minb, unsigned long long* __restrict__ keys,
             unsigned long long* __restrict__ sums, int* __restrict__ counts, unsigned long long mask,
             int* __restrict__ count_err) {
  const double half = __dmul_rn(voxel, 0.5);
  const double o0 = __dsub_rn(vox_unordered(minb[0]), half);
  const double o1 = __dsub_rn(vox_unordered(minb[1]), half);
  const double o2 = __dsub_rn(vox_unordered(minb[2]), half);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double r0 = __dsub_rn(pts[i * 3 + 0], o0);
    const double r1 = __dsub_rn(pts[i * 3 + 1], o1);
    const double r2 = __dsub_rn(pts[i * 3 + 2], o2);
    const double f0 = floor(__ddiv_rn(r0, voxel));
    const double f1 = floor(__ddiv_rn(r1, voxel));
    const double f2 = floor(__ddiv_rn(r2, voxel));
    if (!(f0 >= 0.0 && f0 < kVoxMaxIndex && f1 >= 0.0 && f1 < kVoxMaxIndex && f2 >= 0.0 &&
          f2 < kVoxMaxIndex)) {           // NaN / inf coordinates or more than 2^21 voxels per axis
      atomicExch(count_err + 1, 1);
      continue;
    }
    const unsigned long long key = ((unsigned long long)f0 << 42) | ((unsigned long long)f1 << 21) |
                                   (unsigned long long)f2;
    unsigned long long slot = vox_mix(key) & mask;
    bool placed = false;                  // the table has >= 2 slots per point, so a free slot exists;
    for (unsigned long long probe = 0; probe <= mask; ++probe) {   // the bound only guards against hangs
      const unsigned long long prev = atomicCAS(keys + slot, kVoxEmpty, key);
      if (prev == kVoxEmpty || prev == key) { placed = true; break; }
      slot = (slot + 1) & mask;
    }
    if (!placed) {
      atomicExch(count_err + 1, 1);
      continue;
    }
    atomicAdd(sums + slot * 3 + 0, (unsigned long long)__double2ll_rn(__dmul_rn(r0, kVoxFix)));
    atomicAdd(sums + slot * 3 + 1, (unsigned long long)__double2ll_rn(__dmul_rn(r1, kVoxFix)));
    atomicAdd(sums + slot * 3 + 2, (unsigned long long)__double2ll_rn(__dmul_rn(r2, kVoxFix)));
    atomicAdd(counts + slot, 1);
  }
}

__global__ void __launch_bounds__(256)
k_vox_emit(const unsigned long long* __restrict__ minb, const unsigned long long* __restrict__ keys,
           const unsigned long long* __restrict__ sums, const int* __restrict__ counts,
           unsigned long long cap, double voxel, double* __restrict__ centroids,
           long long* __restrict__ keys_out, int* __restrict__ count_err) {
  const double half = __dmul_rn(voxel, 0.5);
  for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < cap;
       s += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long key = keys[s];
    if (key == kVoxEmpty) continue;
    const int pos = atomicAdd(count_err, 1);
    keys_out[pos] = (long long)key;
    const double den = __dmul_rn((double)counts[s], kVoxFix);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double origin = __dsub_rn(vox_unordered(minb[a]), half);
      const double mean = __ddiv_rn((double)(long long)sums[s * 3 + a], den);
      centroids[(size_t)pos * 3 + a] = __dadd_rn(origin, mean);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Overlap ratio of two clouds (generate_gt.py:68-102): the fraction of query points that
// have a target point closer than `radius` (squared distance < radius^2, the strict test of the
// nanoflann radius search behind open3d's KDTreeFlann.search_radius_vector_3d).  The targets go into a
// hash grid of cell size `radius` (one singly linked list per cell, built with atomicExch), a query
// looks at its 27 neighbouring cells and stops at the first hit.  The count is exact and does not
// depend on the order of arrival.
// ------------------------------------------------------------------------------------------------
constexpr double kOvlBias = 1048576.0;          // 2^20: cell indices are offset to be positive

__device__ __forceinline__ bool ovl_cell(const double* __restrict__ p, double radius, long long c[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double f = __dadd_rn(floor(__ddiv_rn(p[a], radius)), kOvlBias);
    if (!(f >= 1.0 && f < kVoxMaxIndex - 1.0)) return false;   // non-finite or outside +-2^20 cells
    c[a] = (long long)f;
  }
  return true;
}
__device__ __forceinline__ unsigned long long ovl_key(long long x, long long y, long long z) {
  return ((unsigned long long)x << 42) | ((unsigned long long)y << 21) | (unsigned long long)z;
}

__global__ void __launch_bounds__(256)
k_ovl_build(const double* __restrict__ tgt, long long n, double radius,
            unsigned long long* __restrict__ keys, int* __restrict__ head, int* __restrict__ next,
            unsigned long long mask, int* __restrict__ count_err) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    long long c[3];
    if (!ovl_cell(tgt + i * 3, radius, c)) {
      next[i] = -1;
      atomicExch(count_err + 1, 1);
      continue;
    }
    const unsigned long long key = ovl_key(c[0], c[1], c[2]);
    unsigned long long slot = vox_mix(key) & mask;
    bool placed = false;
    for (unsigned long long probe = 0; probe <= mask; ++probe) {   // bounded: never spins on a full table
      const unsigned long long prev = atomicCAS(keys + slot, kVoxEmpty, key);
      if (prev == kVoxEmpty || prev == key) { placed = true; break; }
      slot = (slot + 1) & mask;
    }
    if (!placed) {
      next[i] = -1;
      atomicExch(count_err + 1, 1);
      continue;
    }
    next[i] = atomicExch(head + slot, (int)i);      // push front
  }
}

__global__ void __launch_bounds__(256)
k_ovl_query(const double* __restrict__ qry, long long nq, const double* __restrict__ tgt, double radius,
            const unsigned long long* __restrict__ keys, const int* __restrict__ head,
            const int* __restrict__ next, unsigned long long mask, int* __restrict__ count_err) {
  const double r2 = __dmul_rn(radius, radius);
  int hits = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nq;
       i += (long long)gridDim.x * blockDim.x) {
    const double q0 = qry[i * 3 + 0], q1 = qry[i * 3 + 1], q2 = qry[i * 3 + 2];
    long long c[3];
    if (!ovl_cell(qry + i * 3, radius, c)) {
      atomicExch(count_err + 1, 1);
      continue;
    }
    bool found = false;
    for (int d = 0; d < 27 && !found; ++d) {
      const unsigned long long key = ovl_key(c[0] + d / 9 - 1, c[1] + (d / 3) % 3 - 1, c[2] + d % 3 - 1);
      unsigned long long slot = vox_mix(key) & mask;
      for (unsigned long long probe = 0; probe <= mask; ++probe) {
        const unsigned long long k = keys[slot];
        if (k == kVoxEmpty) break;
        if (k == key) {
          int walked = 0;                 // a list cannot be longer than the table; guards against cycles
          for (int j = head[slot]; j >= 0 && !found && walked <= (int)mask; j = next[j], ++walked) {
            const double dx = __dsub_rn(tgt[(size_t)j * 3 + 0], q0);
            const double dy = __dsub_rn(tgt[(size_t)j * 3 + 1], q1);
            const double dz = __dsub_rn(tgt[(size_t)j * 3 + 2], q2);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            found = d2 < r2;
          }
          break;
        }
        slot = (slot + 1) & mask;
      }
    }
    hits += found ? 1 : 0;
  }
  if (hits) atomicAdd(count_err, hits);
}

static unsigned long long vox_capacity(long long n) {
  unsigned long long cap = 1024;
  while (cap < 2ull * (unsigned long long)n) cap <<= 1;
  return cap;
}

}  // namespace prg

using namespace prg;

extern "C" __attribute__((visibility("default"))) size_t prg_voxel_downsample_workspace_bytes(int64_t n_points) {
  if (n_points < 0) return 0;
  const unsigned long long cap = vox_capacity(n_points);
  return 32 + cap * (8 + 24 + 4);
}

extern "C" __attribute__((visibility("default"))) int prg_voxel_downsample_f64(
    const double* points, int64_t n_points, double voxel_size, double* centroids, int64_t* keys_out,
    int32_t* count_err, void* workspace, size_t workspace_bytes, prg_stream_t stream) {
  PRG_CHECK_ARG(count_err != nullptr, "null pointer");
  PtrDeviceGuard guard(count_err);
  cudaStream_t s = (cudaStream_t)stream;
  PRG_CUDA_OK(cudaMemsetAsync(count_err, 0, 2 * sizeof(int32_t), s));
  if (n_points == 0) return PRG_OK;
  PRG_CHECK_ARG(points && centroids && keys_out && workspace, "null pointer");
  PRG_CHECK_ARG(n_points > 0 && n_points <= (1ll << 30), "point count");
  PRG_CHECK_ARG(voxel_size > 0.0, "voxel size must be positive");
  const unsigned long long cap = vox_capacity(n_points);
  PRG_CHECK_ARG(workspace_bytes >= 32 + cap * 36, "workspace too small");
  PRG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  // [min bounds 4 x u64 | keys cap x u64]  filled with 0xFF;  [sums 3 cap x u64 | counts cap x i32]  zeroed
  unsigned long long* minb = static_cast<unsigned long long*>(workspace);
  unsigned long long* keys = minb + 4;
  unsigned long long* sums = keys + cap;
  int* counts = reinterpret_cast<int*>(sums + 3 * cap);
  PRG_CUDA_OK(cudaMemsetAsync(minb, 0xFF, 32 + cap * 8, s));
  PRG_CUDA_OK(cudaMemsetAsync(sums, 0, cap * 28, s));
  int blocks = ceil_div(n_points, 256 * 4);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  k_vox_bounds<<<blocks, 256, 0, s>>>(points, (long long)n_points, minb);
  PRG_LAUNCH_CHECK();
  k_vox_insert<<<blocks, 256, 0, s>>>(points, (long long)n_points, voxel_size, minb, keys, sums, counts,
                                      cap - 1, count_err);
  PRG_LAUNCH_CHECK();
  int eblocks = ceil_div((int64_t)cap, 256 * 4);
  if (eblocks > num_sms() * 8) eblocks = num_sms() * 8;
  k_vox_emit<<<eblocks, 256, 0, s>>>(minb, keys, sums, counts, cap, voxel_size, centroids,
                                     reinterpret_cast<long long*>(keys_out), count_err);
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}

extern "C" __attribute__((visibility("default"))) size_t prg_overlap_workspace_bytes(int64_t n_target) {
  if (n_target < 0) return 0;
  return vox_capacity(n_target) * 12 + (size_t)n_target * 4;
}

extern "C" __attribute__((visibility("default"))) int prg_overlap_count_f64(
    const double* query, int64_t n_query, const double* target, int64_t n_target, double radius,
    int32_t* count_err, void* workspace, size_t workspace_bytes, prg_stream_t stream) {
  PRG_CHECK_ARG(count_err != nullptr, "null pointer");
  PtrDeviceGuard guard(count_err);
  cudaStream_t s = (cudaStream_t)stream;
  PRG_CUDA_OK(cudaMemsetAsync(count_err, 0, 2 * sizeof(int32_t), s));
  if (n_query == 0 || n_target == 0) return PRG_OK;
  PRG_CHECK_ARG(query && target && workspace, "null pointer");
  PRG_CHECK_ARG(n_query > 0 && n_target > 0 && n_query <= (1ll << 30) && n_target <= (1ll << 30), "point count");
  PRG_CHECK_ARG(radius > 0.0, "radius must be positive");
  const unsigned long long cap = vox_capacity(n_target);
  PRG_CHECK_ARG(workspace_bytes >= cap * 12 + (size_t)n_target * 4, "workspace too small");
  PRG_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  // [keys cap x u64 | head cap x i32] filled with 0xFF (empty key, head = -1); [next n_target x i32]
  unsigned long long* keys = static_cast<unsigned long long*>(workspace);
  int* head = reinterpret_cast<int*>(keys + cap);
  int* next = head + cap;
  PRG_CUDA_OK(cudaMemsetAsync(keys, 0xFF, cap * 12, s));
  int blocks = ceil_div(n_target, 256 * 4);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  k_ovl_build<<<blocks, 256, 0, s>>>(target, (long long)n_target, radius, keys, head, next, cap - 1, count_err);
  PRG_LAUNCH_CHECK();
  blocks = ceil_div(n_query, 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  k_ovl_query<<<blocks, 256, 0, s>>>(query, (long long)n_query, target, radius, keys, head, next, cap - 1,
                                     count_err);
  PRG_LAUNCH_CHECK();
  return PRG_OK;
}
