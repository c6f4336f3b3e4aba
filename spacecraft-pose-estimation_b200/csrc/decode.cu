// Heatmap decode for sm_100a: argmax + quarter-pixel refine + inverse box affine.
//
// Replaces get_max_preds / get_final_preds / transform_preds of the reference
// (landmark_regression/lib/core/inference.py:18-79, lib/utils/transforms.py:49-110); exact
// semantics are restated in SURVEY.md App. A and checked against the reference's own outputs in
// tests/test_decode_gpu.py.
//
// HBM-bound: one streaming read of B*J*H*W floats, O(1) flop per byte.  Design:
//   * persistent CTAs (one per SM), one warp per (frame, landmark) heatmap, maps dealt
//     round-robin to warps so the chip walks HBM as one moving window
//   * every warp owns a ring of shared-memory stages filled by bulk async copies
//     (cp.async.bulk global->shared, SASS UBLKCP) that complete on an mbarrier; lane 0 is the
//     producer, the whole warp is the consumer, so there is no CTA-wide barrier in the loop and
//     the bytes in flight per SM (warps x stages x chunk) do not depend on register pressure
//   * the consumer reads its stage with conflict-free 128-bit shared loads into four independent
//     (max, first-index) accumulators per lane; the warp-wide result takes two REDUX instructions
//   * maps that fit one stage take their refinement neighbours from shared memory, longer maps
//     read the four neighbours back from L2
//   * the per-map epilogue (index -> x/y, mask, refine, FP64 inverse affine, stores) is batched:
//     results are parked in shared memory and finished 32 maps at a time, one map per lane
//   * a plain coalesced-load kernel covers shapes whose maps are not 16-byte aligned
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>
#include <new>
#include <unordered_map>

#include "decode.cuh"
#include "device_util.cuh"

namespace spe {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kNoIndex = 0x7fffffff;

// ------------------------------------------------------------------------------------------
// (value, first index) bookkeeping with NumPy's argmax semantics: strict > keeps the first
// maximum, +-0 compare equal, NaN is sticky in the value (max.NaN) and resolved afterwards.
struct Best {
  float v;
  int i;
};

__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

__device__ __forceinline__ void take(Best& b, float x, int idx) {
  b.i = (x > b.v) ? idx : b.i;  // ordered compare: false as soon as a NaN is involved
  b.v = max_nan(b.v, x);
}

__device__ __forceinline__ void merge(Best& a, float v, int i) {
  // a and (v, i) are both "first maximum of a subset": larger value wins, ties -> lower index
  const bool better = (v > a.v) | ((v == a.v) & (i < a.i));
  a.i = better ? i : a.i;
  a.v = max_nan(a.v, v);
}

// Warp-wide (max, first index) with two REDUX instructions.  Floats are mapped to unsigned keys
// that sort like the values (-0 folded into +0 so that +-0 tie on index, as in NumPy).  A NaN
// anywhere makes the result NaN; its index is resolved by first_nan_index().
__device__ __forceinline__ Best warp_merge(Best b) {
  const bool any_nan = __any_sync(kFull, b.v != b.v);
  const unsigned u = __float_as_uint(b.v + 0.0f);
  const unsigned key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  const unsigned kmax = __reduce_max_sync(kFull, key);
  const int imin = __reduce_min_sync(kFull, key == kmax ? b.i : kNoIndex);
  const unsigned back = (kmax & 0x80000000u) ? (kmax & 0x7fffffffu) : ~kmax;
  return Best{any_nan ? __int_as_float(0x7fc00000) : __uint_as_float(back), imin};
}

// first NaN of a map (only reached when the maximum is NaN)
__device__ __noinline__ int first_nan_index(const float* __restrict__ map, int hw, int lane) {
  for (int base = 0; base < hw; base += 32) {
    const int e = base + lane;
    const bool is_nan = (e < hw) && (map[e] != map[e]);
    const unsigned m = __ballot_sync(kFull, is_nan);
    if (m) return base + __ffs(m) - 1;
  }
  return 0;
}

__device__ __forceinline__ float quarter_sign(float d) {
  // 0.25 * np.sign(d): sign(0) = 0, sign(NaN) = NaN
  return d > 0.f ? 0.25f : (d < 0.f ? -0.25f : d);
}

// One map's epilogue, executed by ONE lane: mask, refine, inverse affine, stores.
// nb = (left, right, up, down) neighbours of the maximum (only read when the refine applies).
template <typename NeighbourFn>
__device__ __forceinline__ void finish_map(const DecodeArgs& a, int map, float v, int idx, NeighbourFn nb) {
  const int py = idx / a.W, px = idx - py * a.W;
  float x = 0.f, y = 0.f;
  if (v > 0.f) {  // inference.py:42-45 (false for NaN)
    x = (float)px;
    y = (float)py;
    if (a.post_process && px > 1 && px < a.W - 1 && py > 1 && py < a.H - 1) {  // inference.py:62
      float l, r, u, d;
      nb(l, r, u, d);
      x += quarter_sign(__fsub_rn(r, l));
      y += quarter_sign(__fsub_rn(d, u));
    }
  }
  if (a.center != nullptr) {
    // Inverse box affine, replaying the float32 roundings of get_affine_transform
    // (transforms.py:65-85).  scale[:,1] never enters.
    const int f = map / a.J;
    const float2 c = __ldg(reinterpret_cast<const float2*>(a.center) + f);
    const float sw = __fmul_rn(__ldg(a.scale + 2 * f), 200.0f);
    const float q1y = __fsub_rn(c.y, __fmul_rn(sw, 0.5f));
    const float dd = __fsub_rn(c.y, q1y);
    const float q2x = __fsub_rn(c.x, dd);
    const double half_w = 0.5 * (double)a.W, half_h = 0.5 * (double)a.H;
    const double ax = __ddiv_rn(__dsub_rn((double)c.x, (double)q2x), half_w);
    const double ay = __ddiv_rn(__dsub_rn((double)c.y, (double)q1y), half_w);
    const double bx = __dsub_rn((double)c.x, __dmul_rn(ax, half_w));
    const double by = __dsub_rn((double)c.y, __dmul_rn(ay, half_h));
    x = (float)__dadd_rn(__dmul_rn(ax, (double)x), bx);
    y = (float)__dadd_rn(__dmul_rn(ay, (double)y), by);
  }
  if (a.kpts != nullptr) {
    float* o = a.kpts + 3 * (size_t)map;
    o[0] = x;
    o[1] = y;
    o[2] = v;
  } else {
    *reinterpret_cast<float2*>(a.preds + 2 * (size_t)map) = make_float2(x, y);
    a.maxvals[map] = v;
  }
  if (a.argmax != nullptr) a.argmax[map] = idx;
}

// ------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ------------------------------------------------------------------------------------------
// Bulk-copy kernel.  Requires (H*W) % 4 == 0 and a 16-byte aligned `hm`.
//   kWarps    warps per CTA (one CTA per SM)
//   kStages   ring depth per warp
//   kChunk    floats per stage (multiple of 128)
struct Pending {  // a decoded map waiting for its epilogue
  float v;
  int idx;
  float nb[4];  // left, right, up, down (single-stage maps only)
};

template <int kWarps, int kStages, int kChunk, int kBatch>
struct BulkLayout {
  static constexpr size_t ring_bytes = (size_t)kWarps * kStages * kChunk * sizeof(float);
  static constexpr size_t pend_bytes = (size_t)kWarps * kBatch * sizeof(Pending);
  static constexpr size_t bar_bytes = (size_t)kWarps * kStages * sizeof(uint64_t);
  static constexpr size_t total = ring_bytes + pend_bytes + bar_bytes;
};

template <int kWarps, int kStages, int kChunk, int kBatch>
__global__ void __launch_bounds__(kWarps * 32, 1) decode_bulk_kernel(const DecodeArgs a) {
  static_assert(kBatch <= 32, "one pending map per lane");
  using Layout = BulkLayout<kWarps, kStages, kChunk, kBatch>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(smem_raw) + (size_t)warp * kStages * kChunk;
  Pending* pend = reinterpret_cast<Pending*>(smem_raw + Layout::ring_bytes) + warp * kBatch;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + Layout::ring_bytes + Layout::pend_bytes) + warp * kStages;

  const int hw = a.H * a.W;
  const int chunks_per_map = (hw + kChunk - 1) / kChunk;
  const bool single = chunks_per_map == 1;
  const int n_warps = gridDim.x * kWarps;
  const int gwarp = blockIdx.x * kWarps + warp;
  // maps gwarp, gwarp + n_warps, ... -> a flat sequence of chunks for this warp
  const int my_maps = (a.n_maps > gwarp) ? (a.n_maps - gwarp + n_warps - 1) / n_warps : 0;
  const long long my_chunks = (long long)my_maps * chunks_per_map;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const uint64_t policy = evict_first_policy();
  auto issue = [&](long long c) {  // lane 0 only
    const int mi = (int)(c / chunks_per_map), ci = (int)(c - (long long)mi * chunks_per_map);
    const int map = gwarp + mi * n_warps;
    const int off = ci * kChunk;
    const int n = min(kChunk, hw - off);
    const int s = (int)(c % kStages);
    const uint32_t bar = smem_u32(&bars[s]);
    mbar_expect_tx(bar, (uint32_t)n * 4u);
    bulk_g2s(smem_u32(ring + (size_t)s * kChunk), a.hm + (size_t)map * hw + off, (uint32_t)n * 4u, bar, policy);
  };
  if (lane == 0) {
    for (long long c = 0; c < my_chunks && c < kStages; ++c) issue(c);
  }

  // epilogue of the maps parked in `pend`: entry e belongs to map (first_mi + e)
  auto flush = [&](int first_mi, int count) {
    __syncwarp();
    if (lane < count) {
      const Pending p = pend[lane];
      const int map = gwarp + (first_mi + lane) * n_warps;
      if (single) {
        finish_map(a, map, p.v, p.idx, [&](float& l, float& r, float& u, float& d) {
          l = p.nb[0], r = p.nb[1], u = p.nb[2], d = p.nb[3];
        });
      } else {
        const float* g = a.hm + (size_t)map * hw + p.idx;
        finish_map(a, map, p.v, p.idx, [&](float& l, float& r, float& u, float& d) {
          l = __ldg(g - 1), r = __ldg(g + 1), u = __ldg(g - a.W), d = __ldg(g + a.W);
        });
      }
    }
    __syncwarp();
  };

  Best acc[4];
  int n_pend = 0, pend_first = 0;
  int mi = 0, ci = 0;
  for (long long c = 0; c < my_chunks; ++c) {
    const int map = gwarp + mi * n_warps;
    const int off = ci * kChunk;
    const int n = min(kChunk, hw - off);
    const int s = (int)(c % kStages);
    const float* stage = ring + (size_t)s * kChunk;
    if (ci == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = Best{-INFINITY, kNoIndex};
    }
    mbar_wait(smem_u32(&bars[s]), (uint32_t)((c / kStages) & 1));

    const int nvec = n >> 2;
    const float4* v4 = reinterpret_cast<const float4*>(stage);
    int ebase = off + 4 * lane;
#pragma unroll 4
    for (int v = lane; v < nvec; v += 32, ebase += 128) {
      const float4 q = v4[v];
      take(acc[0], q.x, ebase);
      take(acc[1], q.y, ebase);
      take(acc[2], q.z, ebase);
      take(acc[3], q.w, ebase);
    }

    if (ci == chunks_per_map - 1) {
      // A lane that never saw a value above -inf still owns its first element (all -inf maps).
      const bool owns = 4 * lane < hw;
      Best b{acc[0].v, acc[0].i == kNoIndex ? 4 * lane : acc[0].i};
#pragma unroll
      for (int k = 1; k < 4; ++k) merge(b, acc[k].v, (acc[k].i == kNoIndex ? 4 * lane : acc[k].i) + k);
      if (!owns) b = Best{-INFINITY, kNoIndex};
      b = warp_merge(b);
      if (b.v != b.v) b.i = first_nan_index(a.hm + (size_t)map * hw, hw, lane);
      if (n_pend == 0) pend_first = mi;
      if (lane == 0) {
        pend[n_pend].v = b.v;
        pend[n_pend].idx = b.i;
      }
      if (single && lane < 4) {
        // clamped so the read stays inside the stage; finish_map ignores it when out of range
        const int d = (lane == 0) ? -1 : (lane == 1) ? 1 : (lane == 2) ? -a.W : a.W;
        pend[n_pend].nb[lane] = stage[min(max(b.i + d, 0), hw - 1)];
      }
      ++n_pend;
    }
    __syncwarp();  // every lane is done with this stage
    if (lane == 0 && c + kStages < my_chunks) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before async overwrite
      issue(c + kStages);
    }
    if (n_pend == kBatch) {
      flush(pend_first, kBatch);
      n_pend = 0;
    }
    if (++ci == chunks_per_map) {
      ci = 0;
      ++mi;
    }
  }
  if (n_pend > 0) flush(pend_first, n_pend);
}

// ------------------------------------------------------------------------------------------
// Dynamically scheduled variant of the above (the default): same per-warp stage, bulk copy, consume
// and batched epilogue, but a warp CLAIMS its next map from a global counter instead of owning the
// fixed sequence gwarp, gwarp + n_warps, ...  A persistent kernel with a static split finishes when
// its slowest CTA does; when part of the chip is busy with another kernel (the previous batch's refit
// tail holds whole SMs for ~0.2 ms, an HRNet forward would do the same) the CTAs that have to wait for
// an SM arrive to find their share already decoded by the others.  Costs nothing when the kernel has
// the chip to itself: the claim for the map after next is issued right behind the bulk copy of the
// next one, so its L2 round trip hides under the copy.
//   counter protocol: a warp claims kDecodeClaim consecutive maps per atomicAdd until it draws an index >= n_maps.
//   The counter comes from a per-(device, stream) slot of two counters used alternately; each launch first zeroes
//   the one the next launch will use (dyn_counter / launch_dyn below).
struct PendingDyn {
  float v;
  int idx;
  float nb[4];
  int map;
  int pad_;
};

template <int kWarps, int kChunk, int kBatch>
struct DynLayout {
  static constexpr size_t ring_bytes = (size_t)kWarps * kChunk * sizeof(float);
  static constexpr size_t pend_bytes = (size_t)kWarps * kBatch * sizeof(PendingDyn);
  static constexpr size_t bar_bytes = (size_t)kWarps * sizeof(uint64_t);
  static constexpr size_t total = ring_bytes + pend_bytes + bar_bytes;
};

#ifndef SPE_DECODE_CLAIM
#define SPE_DECODE_CLAIM 2
#endif
constexpr int kDecodeClaim = SPE_DECODE_CLAIM;

template <int kWarps, int kChunk, int kBatch>
__global__ void __launch_bounds__(kWarps * 32, 1) decode_dyn_kernel(const DecodeArgs a, unsigned* __restrict__ counter, unsigned* __restrict__ next_counter) {
  static_assert(kBatch <= 32, "one pending map per lane");
  using Layout = DynLayout<kWarps, kChunk, kBatch>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* stage = reinterpret_cast<float*>(smem_raw) + (size_t)warp * kChunk;
  PendingDyn* pend = reinterpret_cast<PendingDyn*>(smem_raw + Layout::ring_bytes) + warp * kBatch;
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem_raw + Layout::ring_bytes + Layout::pend_bytes) + warp;
  const uint32_t bar = smem_u32(bar_ptr);

  const int hw = a.H * a.W;
  const int chunks_per_map = (hw + kChunk - 1) / kChunk;
  const bool single = chunks_per_map == 1;

  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // counter[0] = claims, counter[1] = CTAs that have finished.  The last CTA to finish puts both back to zero, so a
  // completed launch always leaves its slot clean (a captured graph may replay the same launch any number of times);
  // in addition every launch first clears the slot the NEXT launch on this stream will use (launch_dyn), so that a
  // launch that did not run to completion cannot poison later ones.
  if (blockIdx.x == 0 && threadIdx.x == 0) next_counter[0] = 0u, next_counter[1] = 0u;
  __syncwarp();
  const uint64_t policy = evict_first_policy();
  // kDecodeClaim consecutive maps per atomic: all warps draw from ONE address, and the rate of same-address atomics
  // (2.4 - 3.4 ns each depending on the state the line is in) is within reach of the kernel's own duration when every
  // map costs one (45 k maps: 0.11 - 0.155 ms against 0.105 ms of HBM time; measured, profiles/decode_alone_r2.md)
  int claim_base = 0, claim_left = 0;
  auto claim = [&]() -> int {  // lane 0 only; >= n_maps: nothing left
    if (claim_left == 0) {
      claim_base = (int)min(atomicAdd(counter, (unsigned)kDecodeClaim), 0x7fffff00u);
      claim_left = kDecodeClaim;
    }
    --claim_left;
    return claim_base++;
  };
  auto issue = [&](int map, int ci) {  // lane 0 only
    const int off = ci * kChunk;
    const int n = min(kChunk, hw - off);
    mbar_expect_tx(bar, (uint32_t)n * 4u);
    bulk_g2s(smem_u32(stage), a.hm + (size_t)map * hw + off, (uint32_t)n * 4u, bar, policy);
  };
  auto flush = [&](int count) {
    __syncwarp();
    if (lane < count) {
      const PendingDyn p = pend[lane];
      if (single) {
        finish_map(a, p.map, p.v, p.idx, [&](float& l, float& r, float& u, float& d) { l = p.nb[0], r = p.nb[1], u = p.nb[2], d = p.nb[3]; });
      } else {
        const float* g = a.hm + (size_t)p.map * hw + p.idx;
        finish_map(a, p.map, p.v, p.idx, [&](float& l, float& r, float& u, float& d) {
          l = __ldg(g - 1), r = __ldg(g + 1), u = __ldg(g - a.W), d = __ldg(g + a.W);
        });
      }
    }
    __syncwarp();
  };

  // lane 0 holds the schedule: the map in flight and the one claimed for afterwards
  int cur = a.n_maps, nxt = a.n_maps;
  if (lane == 0) {
    cur = claim();
    if (cur < a.n_maps) {
      issue(cur, 0);
      nxt = claim();
    }
  }
  cur = __shfl_sync(kFull, cur, 0);

  Best acc[4];
  int n_pend = 0;
  unsigned phase = 0;
  while (cur < a.n_maps) {
    for (int ci = 0; ci < chunks_per_map; ++ci) {
      const int off = ci * kChunk;
      const int n = min(kChunk, hw - off);
      if (ci == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = Best{-INFINITY, kNoIndex};
      }
      mbar_wait(bar, phase & 1u);
      ++phase;
      const int nvec = n >> 2;
      const float4* v4 = reinterpret_cast<const float4*>(stage);
      int ebase = off + 4 * lane;
#pragma unroll 4
      for (int v = lane; v < nvec; v += 32, ebase += 128) {
        const float4 q = v4[v];
        take(acc[0], q.x, ebase);
        take(acc[1], q.y, ebase);
        take(acc[2], q.z, ebase);
        take(acc[3], q.w, ebase);
      }
      const bool last = ci == chunks_per_map - 1;
      if (last) {
        const bool owns = 4 * lane < hw;  // a lane that never saw a value above -inf still owns its first element
        Best b{acc[0].v, acc[0].i == kNoIndex ? 4 * lane : acc[0].i};
#pragma unroll
        for (int k = 1; k < 4; ++k) merge(b, acc[k].v, (acc[k].i == kNoIndex ? 4 * lane : acc[k].i) + k);
        if (!owns) b = Best{-INFINITY, kNoIndex};
        b = warp_merge(b);
        if (b.v != b.v) b.i = first_nan_index(a.hm + (size_t)cur * hw, hw, lane);
        if (lane == 0) {
          pend[n_pend].v = b.v;
          pend[n_pend].idx = b.i;
          pend[n_pend].map = cur;
        }
        if (single && lane < 4) {
          // clamped so the read stays inside the stage; finish_map ignores it when out of range
          const int d = (lane == 0) ? -1 : (lane == 1) ? 1 : (lane == 2) ? -a.W : a.W;
          pend[n_pend].nb[lane] = stage[min(max(b.i + d, 0), hw - 1)];
        }
        ++n_pend;
      }
      __syncwarp();  // every lane is done with the stage
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before the async overwrite
        if (!last) {
          issue(cur, ci + 1);
        } else {
          cur = nxt;
          if (cur < a.n_maps) {
            issue(cur, 0);
            nxt = claim();
          }
        }
      }
      if (last) cur = __shfl_sync(kFull, cur, 0);
    }
    if (n_pend == kBatch) {
      flush(kBatch);
      n_pend = 0;
    }
  }
  if (n_pend > 0) flush(n_pend);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(counter + 1, 1u) == gridDim.x - 1u) {
      counter[0] = 0u;
      counter[1] = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Team variant for FEW, LARGE maps (the reference's 384x384 / 768x768 heatmaps at small batch):
// with one warp per map a batch of 8 x 11 maps would keep 13 SMs busy.  Here a CTA owns a map, its
// warps stream interleaved 16 KB chunks of it (same bulk-copy + mbarrier mechanics, one stage per
// warp), and the per-warp (max, first index) results meet in shared memory.
template <int kWarps, int kChunk>
__global__ void __launch_bounds__(kWarps * 32, 1) decode_team_kernel(const DecodeArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* stage = reinterpret_cast<float*>(smem_raw) + (size_t)warp * kChunk;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kWarps * kChunk * sizeof(float)) + warp;
  Best* team = reinterpret_cast<Best*>(smem_raw + (size_t)kWarps * kChunk * sizeof(float) + kWarps * sizeof(uint64_t));

  const int hw = a.H * a.W;
  const int chunks_per_map = (hw + kChunk - 1) / kChunk;
  const int my_per_map = (chunks_per_map - warp + kWarps - 1) / kWarps;  // >= 1: the launcher requires chunks_per_map >= kWarps
  const int my_maps = (a.n_maps > (int)blockIdx.x) ? (a.n_maps - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long my_chunks = (long long)my_maps * my_per_map;

  if (lane == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint64_t policy = evict_first_policy();
  auto issue = [&](long long c) {  // lane 0 only
    const int mi = (int)(c / my_per_map), k = (int)(c - (long long)mi * my_per_map);
    const int map = blockIdx.x + mi * gridDim.x;
    const int off = (warp + k * kWarps) * kChunk;
    const int n = min(kChunk, hw - off);
    mbar_expect_tx(smem_u32(bar), (uint32_t)n * 4u);
    bulk_g2s(smem_u32(stage), a.hm + (size_t)map * hw + off, (uint32_t)n * 4u, smem_u32(bar), policy);
  };
  if (lane == 0 && my_chunks > 0) issue(0);

  Best acc[4];
  long long c = 0;
  for (int mi = 0; mi < my_maps; ++mi) {
    const int map = blockIdx.x + mi * gridDim.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = Best{-INFINITY, kNoIndex};
    int first_elem = kNoIndex;  // first element this lane owns in this map (for all -inf maps)
    for (int k = 0; k < my_per_map; ++k, ++c) {
      const int off = (warp + k * kWarps) * kChunk;
      const int n = min(kChunk, hw - off);
      mbar_wait(smem_u32(bar), (uint32_t)(c & 1));
      const int nvec = n >> 2;
      const float4* v4 = reinterpret_cast<const float4*>(stage);
      int ebase = off + 4 * lane;
      if (first_elem == kNoIndex && lane < nvec) first_elem = ebase;
#pragma unroll 4
      for (int v = lane; v < nvec; v += 32, ebase += 128) {
        const float4 q = v4[v];
        take(acc[0], q.x, ebase);
        take(acc[1], q.y, ebase);
        take(acc[2], q.z, ebase);
        take(acc[3], q.w, ebase);
      }
      __syncwarp();
      if (lane == 0 && c + 1 < my_chunks) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(c + 1);
      }
    }
    Best b{-INFINITY, kNoIndex};
    if (first_elem != kNoIndex) {
      b = Best{acc[0].v, acc[0].i == kNoIndex ? first_elem : acc[0].i};
#pragma unroll
      for (int k = 1; k < 4; ++k) merge(b, acc[k].v, (acc[k].i == kNoIndex ? first_elem : acc[k].i) + k);
    }
    b = warp_merge(b);
    if (lane == 0) team[warp] = b;
    __syncthreads();
    if (warp == 0) {
      Best t = lane < kWarps ? team[lane] : Best{-INFINITY, kNoIndex};
      t = warp_merge(t);
      const float* g = a.hm + (size_t)map * hw;
      if (t.v != t.v) t.i = first_nan_index(g, hw, lane);
      if (lane == 0) {
        const float* gi = g + t.i;
        finish_map(a, map, t.v, t.i, [&](float& l, float& r, float& u, float& d) {
          l = __ldg(gi - 1), r = __ldg(gi + 1), u = __ldg(gi - a.W), d = __ldg(gi + a.W);
        });
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Fallback for maps that are not 16-byte aligned: warp per map, coalesced scalar loads.
constexpr int kPlainWarps = 4;

__global__ void __launch_bounds__(kPlainWarps * 32) decode_plain_kernel(const DecodeArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hw = a.H * a.W;
  const int n_warps = gridDim.x * kPlainWarps;
  for (int map = blockIdx.x * kPlainWarps + warp; map < a.n_maps; map += n_warps) {
    const float* g = a.hm + (size_t)map * hw;
    Best b{-INFINITY, kNoIndex};
#pragma unroll 4
    for (int e = lane; e < hw; e += 32) take(b, __ldg(g + e), e);
    if (b.i == kNoIndex && lane < hw) b.i = lane;
    b = warp_merge(b);
    if (b.v != b.v) b.i = first_nan_index(g, hw, lane);
    if (lane == 0) {
      const float* gi = g + b.i;
      finish_map(a, map, b.v, b.i, [&](float& l, float& r, float& u, float& d) {
        l = __ldg(gi - 1), r = __ldg(gi + 1), u = __ldg(gi - a.W), d = __ldg(gi + a.W);
      });
    }
  }
}

// ------------------------------------------------------------------------------------------
// f2: fused pre-combination + decode.  The reference averages heatmap tensors before decoding:
//   * flip test (lib/core/function.py:347-366): (out + shift(flip_back(out_flipped))) * 0.5, with
//     flip_back = reverse W + swap matched joints (lib/utils/transforms.py:15-29) and the 1-px
//     shift of TEST.SHIFT_HEATMAP (columns 1.. take the flipped value to their left);
//   * model ensemble (validate_cv, function.py:525-536): ((o0 + o1) + ...) / K — on the GPU torch
//     evaluates tensor / python-scalar as a multiplication by the float32 reciprocal, which is
//     what is reproduced here so that maxvals stay bit-identical to the reference run.
// Combining in the decode read saves writing the averaged tensor and reading it back (2 of K+2
// passes over HBM).  Warp per map, K coalesced 128-bit streams (the flipped stream is read with
// 32-bit loads: its reversed, shifted window is never 16-byte aligned).
struct CombineSrc {
  const float* p[kMaxCombine];
};

template <int kMode, int kK>
__device__ __forceinline__ float combined_at(const CombineArgs& a, const CombineSrc& s, int idx, float inv_k) {
  if (kMode == kCombineMean) {
    float acc = __ldg(s.p[0] + idx);
#pragma unroll
    for (int k = 1; k < kK; ++k) acc = __fadd_rn(acc, __ldg(s.p[k] + idx));
    return __fmul_rn(acc, inv_k);
  } else {
    const int W = a.out.W;
    const int y = idx / W, x = idx - y * W;
    const int xs = a.shift_heatmap ? (x == 0 ? W - 1 : W - x) : W - 1 - x;
    return __fmul_rn(__fadd_rn(__ldg(s.p[0] + idx), __ldg(s.p[1] + y * W + xs)), 0.5f);
  }
}

template <int kMode, bool kVec, int kK>
__global__ void __launch_bounds__(256) decode_combine_kernel(const CombineArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const DecodeArgs& o = a.out;
  const int hw = o.H * o.W, W = o.W;
  const int n_warps = gridDim.x * 8;
  const float inv_k = 1.0f / (float)a.K;
  for (int map = blockIdx.x * 8 + warp; map < o.n_maps; map += n_warps) {
    CombineSrc s;
    if (kMode == kCombineMean) {
#pragma unroll
      for (int k = 0; k < kK; ++k) s.p[k] = a.src[k] + (size_t)map * hw;
    } else {
      const int b = map / o.J, j = map - b * o.J;
      const int jf = a.flip_perm ? a.flip_perm[j] : j;
      s.p[0] = a.src[0] + (size_t)map * hw;
      s.p[1] = a.src[1] + ((size_t)b * o.J + jf) * hw;
    }
    Best acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = Best{-INFINITY, kNoIndex};
    Best b;
    if (kVec) {
      const int nvec = hw >> 2;
#pragma unroll(kK >= 4 ? 1 : 2)
      for (int v = lane; v < nvec; v += 32) {
        float4 q = __ldg(reinterpret_cast<const float4*>(s.p[0]) + v);
        if (kMode == kCombineMean) {
#pragma unroll
          for (int k = 1; k < kK; ++k) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(s.p[k]) + v);
            q.x = __fadd_rn(q.x, r.x), q.y = __fadd_rn(q.y, r.y), q.z = __fadd_rn(q.z, r.z), q.w = __fadd_rn(q.w, r.w);
          }
          q.x = __fmul_rn(q.x, inv_k), q.y = __fmul_rn(q.y, inv_k), q.z = __fmul_rn(q.z, inv_k), q.w = __fmul_rn(q.w, inv_k);
        } else {
          const int e = 4 * v, y = e / W, x = e - y * W;  // W % 4 == 0: the four elements share a row
          const float* row = s.p[1] + y * W;
          float f0, f1, f2, f3;
          if (a.shift_heatmap) {
            f0 = __ldg(row + (x == 0 ? W - 1 : W - x)), f1 = __ldg(row + W - x - 1), f2 = __ldg(row + W - x - 2), f3 = __ldg(row + W - x - 3);
          } else {
            f0 = __ldg(row + W - 1 - x), f1 = __ldg(row + W - 2 - x), f2 = __ldg(row + W - 3 - x), f3 = __ldg(row + W - 4 - x);
          }
          q.x = __fmul_rn(__fadd_rn(q.x, f0), 0.5f), q.y = __fmul_rn(__fadd_rn(q.y, f1), 0.5f);
          q.z = __fmul_rn(__fadd_rn(q.z, f2), 0.5f), q.w = __fmul_rn(__fadd_rn(q.w, f3), 0.5f);
        }
        const int ebase = 4 * v;
        take(acc[0], q.x, ebase);
        take(acc[1], q.y, ebase);
        take(acc[2], q.z, ebase);
        take(acc[3], q.w, ebase);
      }
      const bool owns = 4 * lane < hw;
      b = Best{acc[0].v, acc[0].i == kNoIndex ? 4 * lane : acc[0].i};
#pragma unroll
      for (int k = 1; k < 4; ++k) merge(b, acc[k].v, (acc[k].i == kNoIndex ? 4 * lane : acc[k].i) + k);
      if (!owns) b = Best{-INFINITY, kNoIndex};
    } else {
      b = Best{-INFINITY, kNoIndex};
      for (int e = lane; e < hw; e += 32) take(b, combined_at<kMode, kK>(a, s, e, inv_k), e);
      if (b.i == kNoIndex && lane < hw) b.i = lane;
    }
    b = warp_merge(b);
    if (b.v != b.v) {  // NaN maximum: first NaN of the COMBINED map
      int first = 0;
      for (int base = 0; base < hw; base += 32) {
        const int e = base + lane;
        const float v = e < hw ? combined_at<kMode, kK>(a, s, e, inv_k) : 0.f;
        const unsigned m = __ballot_sync(kFull, v != v);
        if (m) {
          first = base + __ffs(m) - 1;
          break;
        }
      }
      b.i = first;
    }
    if (lane == 0) {
      finish_map(o, map, b.v, b.i, [&](float& l, float& r, float& u, float& d) {
        l = combined_at<kMode, kK>(a, s, b.i - 1, inv_k), r = combined_at<kMode, kK>(a, s, b.i + 1, inv_k);
        u = combined_at<kMode, kK>(a, s, b.i - W, inv_k), d = combined_at<kMode, kK>(a, s, b.i + W, inv_k);
      });
    }
  }
}

template <int kWarps, int kStages, int kChunk, int kBatch = 32, int kCarveoutPct = -1>
cudaError_t launch_bulk(const DecodeArgs& a, int dev, int num_sms, cudaStream_t stream) {
  constexpr size_t smem = BulkLayout<kWarps, kStages, kChunk, kBatch>::total;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static PerDeviceOnce once;
  cudaError_t e = once.run(dev, [] {
    cudaError_t r = cudaFuncSetAttribute(decode_bulk_kernel<kWarps, kStages, kChunk, kBatch>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (r == cudaSuccess && kCarveoutPct >= 0)
      r = cudaFuncSetAttribute(decode_bulk_kernel<kWarps, kStages, kChunk, kBatch>, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveoutPct);
    return r;
  });
  if (e != cudaSuccess) return e;
  const int ctas_needed = (a.n_maps + kWarps - 1) / kWarps;
  const int grid = ctas_needed < num_sms ? ctas_needed : num_sms;
  decode_bulk_kernel<kWarps, kStages, kChunk, kBatch><<<grid, kWarps * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

// Claim counters for decode_dyn_kernel.  The decode entry points take no workspace, so the library keeps one 128-byte
// slot per (device, stream) it has seen: launches on one stream are ordered and may share a slot, launches on different
// streams never do.  A slot holds TWO counters used alternately: every launch zeroes, as its first action, the counter
// the next launch on that stream will draw from, so no launch depends on an earlier one having run to completion (and
// no memset node sits between two kernels of the stream: that costs an engine switch of ~40 us).  A process that uses
// more than kDynSlots streams per device gets nullptr for the others, and those launches take the statically scheduled
// kernel (same results).
constexpr int kDynSlots = 1024;
struct DynSlot {
  int index;
  unsigned parity;
};
inline cudaError_t dyn_counter(int dev, cudaStream_t stream, unsigned** cur, unsigned** next) {
  static std::mutex mu;
  static unsigned* base[kMaxDevices] = {};
  static std::unordered_map<uintptr_t, DynSlot>* slots[kMaxDevices] = {};
  *cur = *next = nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (base[dev] == nullptr) {
    unsigned* p = nullptr;
    cudaError_t r = cudaMalloc(&p, sizeof(unsigned) * kDynSlots * 32);  // one slot per 128 B line
    if (r == cudaSuccess) r = cudaMemset(p, 0, sizeof(unsigned) * kDynSlots * 32);
    if (r != cudaSuccess) {
      if (p) cudaFree(p);
      return r;
    }
    base[dev] = p;
    slots[dev] = new (std::nothrow) std::unordered_map<uintptr_t, DynSlot>();
  }
  if (slots[dev] == nullptr) return cudaSuccess;  // no table: static schedule
  try {
    auto it = slots[dev]->find(reinterpret_cast<uintptr_t>(stream));
    if (it == slots[dev]->end()) {
      if ((int)slots[dev]->size() >= kDynSlots) return cudaSuccess;
      it = slots[dev]->emplace(reinterpret_cast<uintptr_t>(stream), DynSlot{(int)slots[dev]->size(), 0u}).first;
    }
    unsigned* slot = base[dev] + 32 * it->second.index;
    *cur = slot + 16 * (it->second.parity & 1u);
    *next = slot + 16 * ((it->second.parity + 1u) & 1u);
    it->second.parity ^= 1u;
  } catch (...) {  // allocation failure inside the map: static schedule for this launch
    *cur = *next = nullptr;
  }
  return cudaSuccess;
}

template <int kWarps, int kChunk, int kBatch = 32>
cudaError_t launch_dyn(const DecodeArgs& a, int dev, int num_sms, cudaStream_t stream) {
  constexpr size_t smem = DynLayout<kWarps, kChunk, kBatch>::total;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static PerDeviceOnce once;
  cudaError_t e = once.run(dev, [] {
    cudaError_t r = cudaFuncSetAttribute(decode_dyn_kernel<kWarps, kChunk, kBatch>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(decode_dyn_kernel<kWarps, kChunk, kBatch>, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
    return r;
  });
  if (e != cudaSuccess) return e;
  unsigned *counter = nullptr, *next_counter = nullptr;
  e = dyn_counter(dev, stream, &counter, &next_counter);
  if (e != cudaSuccess) return e;
  if (counter == nullptr) return launch_bulk<kWarps, 1, kChunk, kBatch, kSmemCarveoutPct>(a, dev, num_sms, stream);  // same shape, static split
  const int ctas_needed = (a.n_maps + kWarps - 1) / kWarps;
  const int grid = ctas_needed < num_sms ? ctas_needed : num_sms;
  decode_dyn_kernel<kWarps, kChunk, kBatch><<<grid, kWarps * 32, smem, stream>>>(a, counter, next_counter);
  return cudaGetLastError();
}

}  // namespace

// development builds (-DSPE_DEV): SPE_DECODE_VARIANT picks one of the warps x stages x chunk shapes measured in
// profiles/decode_variants_r1.md; the shipped library reads no environment variable and has the default shape only
static int decode_variant() {
#ifdef SPE_DEV
  static const int variant = [] {
    const char* v = getenv("SPE_DECODE_VARIANT");
    return v ? atoi(v) : 0;
  }();
  return variant;
#else
  return 0;
#endif
}

cudaError_t launch_decode(const DecodeArgs& a, cudaStream_t stream) {
  if (a.n_maps == 0) return cudaSuccess;
  int dev = 0, num_sms = 0;
  const cudaError_t de = current_device(dev, num_sms);
  if (de != cudaSuccess) return de;
  const long long hw = (long long)a.H * a.W;
  const bool aligned = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.hm) & 15u) == 0);
  {
    // few large maps: one CTA per map instead of one warp per map
    constexpr int kTeamWarps = 7, kTeamChunk = 4096;
    const long long chunks_per_map = (hw + kTeamChunk - 1) / kTeamChunk;
    if (aligned && decode_variant() == 0 && chunks_per_map >= kTeamWarps && a.n_maps < num_sms * kTeamWarps) {
      constexpr size_t smem = (size_t)kTeamWarps * kTeamChunk * sizeof(float) + kTeamWarps * (sizeof(uint64_t) + sizeof(Best));
      static PerDeviceOnce once;
      cudaError_t e = once.run(dev, [] {
        cudaError_t r = cudaFuncSetAttribute(decode_team_kernel<kTeamWarps, kTeamChunk>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (r == cudaSuccess)
          r = cudaFuncSetAttribute(decode_team_kernel<kTeamWarps, kTeamChunk>, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
        return r;
      });
      if (e != cudaSuccess) return e;
      const int grid = a.n_maps < num_sms ? a.n_maps : num_sms;
      decode_team_kernel<kTeamWarps, kTeamChunk><<<grid, kTeamWarps * 32, smem, stream>>>(a);
      return cudaGetLastError();
    }
  }
  if (aligned && a.background) {
    // 3 warps x one 16 KB stage: 51 KB of shared memory and 6.9 K registers, i.e. the footprint of ONE of the three
    // hypothesis-kernel CTAs of an SM (39 KB, 21.5 K registers): as those retire, a decode CTA of the next batch
    // takes the slot, streams its share of the maps while the SM's FP32 pipes stay busy with the other two, and
    // leaves.  48 KB in flight per SM is still enough for ~6 TB/s (Little), so the slot is held only briefly.
    return launch_dyn<3, 4096, 16>(a, dev, num_sms, stream);
  }
  if (aligned) {
#ifdef SPE_DEV
    switch (decode_variant()) {
      case 1: return launch_bulk<4, 3, 4096>(a, dev, num_sms, stream);   // 192 KB, 4 warps
      case 2: return launch_bulk<8, 3, 2048>(a, dev, num_sms, stream);   // 192 KB, 8 warps, 8 KB stages
      case 3: return launch_bulk<12, 2, 2048>(a, dev, num_sms, stream);  // 192 KB, 12 warps
      case 4: return launch_bulk<6, 2, 4096>(a, dev, num_sms, stream);   // 192 KB, 6 warps x 2 stages x 16 KB
      case 5: return launch_bulk<7, 2, 4096, 16>(a, dev, num_sms, stream);  // 224 KB: 7 warps x 2 stages x 16 KB
      case 6: return launch_bulk<12, 1, 4096>(a, dev, num_sms, stream);  // 192 KB: 12 warps, one 16 KB stage each
      // measured best on B200 (profiles/decode_variants_r1.md): 7 warps, one 16 KB stage each =
      // 112 KB of bulk copies in flight per SM; 64x64 maps refine from shared memory.  The 132 KB
      // carveout (58 %) is the one the pose kernels use too, so an SM never has to drain to switch
      // its shared-memory/L1 split when the kernels of consecutive batches overlap.
      case 7: return launch_bulk<7, 1, 4096, 32, kSmemCarveoutPct>(a, dev, num_sms, stream);  // same shape, static split of the maps
      default: break;
    }
#endif
    // 7 warps, one 16 KB stage each = 112 KB of bulk copies in flight per SM (measured best on B200,
    // profiles/decode_variants_r1.md), maps claimed dynamically (decode_dyn_kernel).  The 132 KB carveout (58 %) is the
    // one the pose kernels use too, so an SM never has to drain to switch its shared-memory/L1 split when the kernels of
    // consecutive batches overlap.
    return launch_dyn<7, 4096>(a, dev, num_sms, stream);
  }
  const int ctas_needed = (a.n_maps + kPlainWarps - 1) / kPlainWarps;
  const int cap = num_sms * 8;
  decode_plain_kernel<<<ctas_needed < cap ? ctas_needed : cap, kPlainWarps * 32, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_decode_combined(const CombineArgs& a, cudaStream_t stream) {
  const DecodeArgs& o = a.out;
  if (o.n_maps == 0) return cudaSuccess;
  int dev = 0, num_sms = 0;
  const cudaError_t de = current_device(dev, num_sms);
  if (de != cudaSuccess) return de;
  bool vec = ((long long)o.H * o.W) % 4 == 0 && o.W % 4 == 0;
  for (int k = 0; k < a.K; ++k) vec = vec && ((reinterpret_cast<uintptr_t>(a.src[k]) & 15u) == 0);
  const int ctas_needed = (o.n_maps + 7) / 8;
  const int cap = num_sms * 8;  // 8 CTAs x 8 warps per SM: ~64 x K x 2 128-bit loads in flight per SM
  const int grid = ctas_needed < cap ? ctas_needed : cap;
  if (a.mode == kCombineMean) {
    switch (a.K * 2 + (vec ? 1 : 0)) {
#define SPE_MEAN_CASE(K_)                                                                                  \
  case K_ * 2 + 1: decode_combine_kernel<kCombineMean, true, K_><<<grid, 256, 0, stream>>>(a); break;     \
  case K_ * 2: decode_combine_kernel<kCombineMean, false, K_><<<grid, 256, 0, stream>>>(a); break;
      SPE_MEAN_CASE(1) SPE_MEAN_CASE(2) SPE_MEAN_CASE(3) SPE_MEAN_CASE(4) SPE_MEAN_CASE(5) SPE_MEAN_CASE(6) SPE_MEAN_CASE(7) SPE_MEAN_CASE(8)
#undef SPE_MEAN_CASE
      default: return cudaErrorInvalidValue;
    }
  } else {
    if (vec) decode_combine_kernel<kCombineFlip, true, 2><<<grid, 256, 0, stream>>>(a);
    else decode_combine_kernel<kCombineFlip, false, 2><<<grid, 256, 0, stream>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace spe
