// K7: the counts behind pangraph's mash distance (mash.h), all of it streaming integer work:
//   1. sketch    one CTA per tile of 4096 positions (mash_core.h: bases staged with 128-bit loads, rolling k-mers, block
//                prefix / suffix minima, per-position decisions from shared memory); run twice -- a counting pass, a scan of the tile counts, a writing pass -- so
//                that the key array has its exact size and no position-sized scratch exists.  A key is value << seq_bits | seq.
//   2. sort      one radix sort of the 64-bit keys, then the distinct keys (a value once per sequence that holds it).
//   3. incidence values held by two or more sequences get a dense rank; a bitmap row per sequence over those ranks.
//   4. pairs     counts[i][j] = popcount(row_i & row_j), 4 x 4 pairs per warp; the diagonal is a histogram of the distinct keys.
// The reference does this on one host thread: sketch every sequence, sort the (value, position) list, and for every value
// add 1 to each pair of its sequences (mash_distance.rs:16-48) -- the same matrix, element for element, since all counts
// are integers.  The divisions that turn counts into distances (mash_distance.rs:57-64) are done by the caller in IEEE f64.
#include "mash.h"

#include <cub/cub.cuh>

#include <algorithm>
#include <memory>
#include <vector>

#include "mash_core.h"
#include "pgmm_cuda.h"

namespace pgmm {
namespace mash {
namespace {

struct MashTile {
  uint32_t seq, t0;  // sequence and first position of the tile inside it
};

inline size_t tile_smem(int w, int k) { return layout_of(w, k).total; }

// WRITE = false: counts[tile] = flagged positions of the tile.  WRITE = true: their keys, in position order, from
// keys[prefix[tile]] on.
template <bool WRITE>
__global__ void __launch_bounds__(kTileThreads) mash_tile_kernel(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ starts,
                                                                  const MashTile *__restrict__ tiles, int w, int k, int seq_bits,
                                                                  uint64_t *__restrict__ counts, const uint64_t *__restrict__ prefix,
                                                                  uint64_t *__restrict__ keys) {
  extern __shared__ __align__(16) uint8_t mash_smem[];
  const MashTile tl = tiles[blockIdx.x];
  const int tid = threadIdx.x;
  const uint64_t s0 = starts[tl.seq];
  const Tile t = tile_of((int64_t)(starts[tl.seq + 1] - s0), (int64_t)tl.t0, w, k);
  const int ce = cap_ev(w);
  const Layout lay = layout_of(w, k);
  uint64_t *X = (uint64_t *)(mash_smem + lay.x);
  uint16_t *EL = (uint16_t *)(mash_smem + lay.el), *P = (uint16_t *)(mash_smem + lay.p), *S = (uint16_t *)(mash_smem + lay.s);
  uint8_t *F = mash_smem + lay.f, *CDraw = mash_smem + lay.cd;

  // stage the bases with 128-bit loads from the 16-byte boundary at or before the first one needed (`bases` is a cudaMalloc
  // block, so that boundary lies inside it), then turn the bytes into codes in place; CD[i] = code of position c_lo + i
  const uint8_t *src = bases + s0 + t.c_lo;
  const int pre = (int)((uintptr_t)src & 15);
  const uint8_t *CD = CDraw + pre;
  {
    const int n16 = (pre + t.n_codes) / 16;
    const uint4 *src16 = (const uint4 *)(src - pre);
    for (int i = tid; i < n16; i += kTileThreads) ((uint4 *)CDraw)[i] = __ldg(src16 + i);
    for (int i = 16 * n16 + tid; i < pre + t.n_codes; i += kTileThreads) CDraw[i] = src[i - pre];
    for (int i = tid; i < ce; i += kTileThreads) F[i] = 0;
  }
  __syncthreads();
  for (int i = tid; i < pre + t.n_codes; i += kTileThreads) CDraw[i] = (uint8_t)code(CDraw[i]);
  __syncthreads();
  roll(t, w, k, tid, kTileThreads, CD, X, EL);
  __syncthreads();
  scan_blocks(t, w, tid, kTileThreads, X, P, S);
  __syncthreads();
  decide(t, w, k, tid, kTileThreads, X, EL, P, S, F);
  __syncthreads();

  typedef cub::BlockScan<int, kTileThreads> Scan;
  __shared__ typename Scan::TempStorage scan_tmp;
  const int base = (int)(t.t0 - t.e_lo);  // slot of the tile's first position
  const int n_own = (int)(t.t1 - t.t0);
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < kTileItems; ++i) {
    const int j = tid * kTileItems + i;
    cnt += j < n_own && F[base + j];
  }
  int off, total;
  Scan(scan_tmp).ExclusiveSum(cnt, off, total);
  if (!WRITE) {
    if (tid == 0) counts[blockIdx.x] = (uint64_t)total;
    return;
  }
  uint64_t *out = keys + prefix[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kTileItems; ++i) {
    const int j = tid * kTileItems + i;
    if (j < n_own && F[base + j]) out[off++] = X[base + j] << seq_bits | (uint64_t)tl.seq;
  }
}

// per distinct key: one more distinct value for its sequence; is its value held by another sequence too (the keys are
// sorted, the distinct keys of one value are neighbours), and is it the first key of such a value
__global__ void mash_mark_kernel(int64_t n, const uint64_t *__restrict__ ukeys, int seq_bits, uint32_t *__restrict__ diag,
                                 uint8_t *__restrict__ shared, uint32_t *__restrict__ head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = ukeys[i], v = key >> seq_bits;
  atomicAdd(&diag[key & ((1ull << seq_bits) - 1)], 1u);
  const bool same_prev = i > 0 && (ukeys[i - 1] >> seq_bits) == v;
  const bool same_next = i + 1 < n && (ukeys[i + 1] >> seq_bits) == v;
  shared[i] = same_prev || same_next;
  head[i] = !same_prev && same_next;
}

// bit (rank of the value among the shared ones) of the sequence's row
__global__ void mash_bitmap_kernel(int64_t n, const uint64_t *__restrict__ ukeys, int seq_bits, const uint8_t *__restrict__ shared,
                                   const uint32_t *__restrict__ head_incl, uint64_t words, unsigned long long *__restrict__ bitmap) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !shared[i]) return;
  const uint64_t seq = ukeys[i] & ((1ull << seq_bits) - 1);
  const uint32_t vid = head_incl[i] - 1;
  atomicOr(&bitmap[seq * words + (vid >> 6)], 1ull << (vid & 63));
}

// one warp per 4 x 4 block of pairs (rows 4 bi .., 4 bj ..), bj >= bi; lanes stride over the words of the rows
constexpr int kPairWarps = 4;
__global__ void __launch_bounds__(32 * kPairWarps) mash_pair_kernel(int n, uint64_t words, const unsigned long long *__restrict__ bitmap,
                                                                     uint32_t *__restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int bi = blockIdx.y, bj = blockIdx.x * kPairWarps + (threadIdx.x >> 5);
  const int nb = (n + 3) / 4;
  if (bj >= nb || bj < bi) return;
  const unsigned long long *ri[4], *rj[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    ri[a] = bitmap + (uint64_t)min(4 * bi + a, n - 1) * words;  // rows past the end repeat the last one (never written)
    rj[a] = bitmap + (uint64_t)min(4 * bj + a, n - 1) * words;
  }
  uint32_t acc[4][4] = {};
  for (uint64_t x = lane; x < words; x += 32) {
    unsigned long long vi[4], vj[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) vi[a] = ri[a][x], vj[a] = rj[a][x];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] += (uint32_t)__popcll(vi[a] & vj[b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      uint32_t v = acc[a][b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const int i = 4 * bi + a, j = 4 * bj + b;
      if (lane == 0 && i < j && j < n) counts[(uint64_t)i * n + j] = v, counts[(uint64_t)j * n + i] = v;
    }
}

__global__ void mash_diag_kernel(int n, const uint32_t *__restrict__ diag, uint32_t *__restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) counts[(uint64_t)i * n + i] = diag[i];
}

// plain device allocation for a job that runs once per build (the pooled DevBuf would keep gigabytes cached)
template <typename T>
struct Dev {
  T *p = nullptr;
  explicit Dev(size_t n) { PGMM_CUDA(cudaMalloc((void **)&p, std::max<size_t>(1, n) * sizeof(T))); }
  ~Dev() { cudaFree(p); }
  Dev(const Dev &) = delete;
  Dev &operator=(const Dev &) = delete;
};

struct Timer {
  cudaEvent_t a, b;
  cudaStream_t st;
  explicit Timer(cudaStream_t s) : st(s) {
    PGMM_CUDA(cudaEventCreate(&a));
    PGMM_CUDA(cudaEventCreate(&b));
    PGMM_CUDA(cudaEventRecord(a, st));
  }
  double stop() {  // waits for the stage
    float ms = 0;
    PGMM_CUDA(cudaEventRecord(b, st));
    PGMM_CUDA(cudaEventSynchronize(b));
    PGMM_CUDA(cudaEventElapsedTime(&ms, a, b));
    return ms;
  }
  ~Timer() { cudaEventDestroy(a), cudaEventDestroy(b); }
};

struct Stream {
  cudaStream_t s = nullptr;
  Stream() { PGMM_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); }
  ~Stream() { cudaStreamDestroy(s); }
  Stream(const Stream &) = delete;
  Stream &operator=(const Stream &) = delete;
};

inline unsigned blocks_of(int64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

}  // namespace

int shared_counts(const char *const *seqs, const int64_t *lens, int n, int k, int w, uint32_t *counts, Stats *stats) {
  if (k < 1 || k > kMaxK || w < 1 || w > kMaxW) return -1;
  if (n < 1) return -2;
  int seq_bits = 1;
  while (seq_bits < 32 && (1ull << seq_bits) < (uint64_t)n) ++seq_bits;
  if (2 * k + seq_bits > 64) return -3;
  std::vector<uint64_t> starts((size_t)n + 1, 0);
  std::vector<MashTile> tiles;
  for (int i = 0; i < n; ++i) {
    if (lens[i] < 0 || lens[i] >= (1ll << 31)) return -4;
    starts[(size_t)i + 1] = starts[(size_t)i] + (uint64_t)lens[i];
    for (int64_t t0 = 0; t0 < lens[i]; t0 += kTile) tiles.push_back(MashTile{(uint32_t)i, (uint32_t)t0});
  }
  require_device();
  Stats st;
  st.bases = starts[(size_t)n], st.tiles = tiles.size();
  Stream stream_owner;
  cudaStream_t stream = stream_owner.s;
  std::fill(counts, counts + (size_t)n * n, 0u);
  Dev<uint32_t> d_counts((size_t)n * n), d_diag((size_t)n);
  PGMM_CUDA(cudaMemsetAsync(d_counts.p, 0, (size_t)n * n * sizeof(uint32_t), stream));
  PGMM_CUDA(cudaMemsetAsync(d_diag.p, 0, (size_t)n * sizeof(uint32_t), stream));

  // ---- 1. sketch ----
  int64_t M = 0;
  std::unique_ptr<Dev<uint64_t>> d_keys_raw;
  if (!tiles.empty()) {
    Timer tu(stream);
    Dev<uint8_t> d_bases((size_t)st.bases + 16);
    Dev<uint64_t> d_starts((size_t)n + 1), d_tile_cnt(tiles.size() + 1), d_tile_off(tiles.size() + 1);
    Dev<MashTile> d_tiles(tiles.size());
    for (int i = 0; i < n; ++i)
      if (lens[i]) PGMM_CUDA(cudaMemcpyAsync(d_bases.p + starts[(size_t)i], seqs[i], (size_t)lens[i], cudaMemcpyHostToDevice, stream));
    PGMM_CUDA(cudaMemcpyAsync(d_starts.p, starts.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, stream));
    PGMM_CUDA(cudaMemcpyAsync(d_tiles.p, tiles.data(), tiles.size() * sizeof(MashTile), cudaMemcpyHostToDevice, stream));
    PGMM_CUDA(cudaMemsetAsync(d_tile_cnt.p, 0, (tiles.size() + 1) * 8, stream));
    st.upload_ms = tu.stop();

    Timer ts(stream);
    const size_t smem = tile_smem(w, k);
    PGMM_CUDA(cudaFuncSetAttribute(mash_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PGMM_CUDA(cudaFuncSetAttribute(mash_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mash_tile_kernel<false><<<(unsigned)tiles.size(), kTileThreads, smem, stream>>>(d_bases.p, d_starts.p, d_tiles.p, w, k, seq_bits,
                                                                                     d_tile_cnt.p, nullptr, nullptr);
    PGMM_CUDA(cudaGetLastError());
    {  // offsets of the tiles' keys; the extra (zero) element makes the last offset the total
      size_t bytes = 0;
      PGMM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_tile_cnt.p, d_tile_off.p, (int64_t)tiles.size() + 1, stream));
      Dev<uint8_t> tmp(bytes);
      PGMM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, d_tile_cnt.p, d_tile_off.p, (int64_t)tiles.size() + 1, stream));
      uint64_t total = 0;
      PGMM_CUDA(cudaMemcpyAsync(&total, d_tile_off.p + tiles.size(), 8, cudaMemcpyDeviceToHost, stream));
      PGMM_CUDA(cudaStreamSynchronize(stream));
      M = (int64_t)total;
    }
    d_keys_raw.reset(new Dev<uint64_t>((size_t)M));
    if (M > 0) {
      mash_tile_kernel<true><<<(unsigned)tiles.size(), kTileThreads, smem, stream>>>(d_bases.p, d_starts.p, d_tiles.p, w, k, seq_bits,
                                                                                      nullptr, d_tile_off.p, d_keys_raw->p);
      PGMM_CUDA(cudaGetLastError());
    }
    st.sketch_ms = ts.stop();
    st.launches += 3 + (M > 0);
  }
  st.minimizers = (uint64_t)M;

  if (M > 0) {
    // ---- 2. sort, distinct keys ----
    Timer tsort(stream);
    Dev<uint64_t> d_sorted((size_t)M), d_ukeys((size_t)M);
    Dev<int64_t> d_nsel(1);
    {
      size_t bytes = 0;
      PGMM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, d_keys_raw->p, d_sorted.p, M, 0, 2 * k + seq_bits, stream));
      Dev<uint8_t> tmp(bytes);
      PGMM_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, d_keys_raw->p, d_sorted.p, M, 0, 2 * k + seq_bits, stream));
      PGMM_CUDA(cudaStreamSynchronize(stream));
    }
    d_keys_raw.reset();
    int64_t U = 0;
    {
      size_t bytes = 0;
      PGMM_CUDA(cub::DeviceSelect::Unique(nullptr, bytes, d_sorted.p, d_ukeys.p, d_nsel.p, M, stream));
      Dev<uint8_t> tmp(bytes);
      PGMM_CUDA(cub::DeviceSelect::Unique(tmp.p, bytes, d_sorted.p, d_ukeys.p, d_nsel.p, M, stream));
      PGMM_CUDA(cudaMemcpyAsync(&U, d_nsel.p, 8, cudaMemcpyDeviceToHost, stream));
      PGMM_CUDA(cudaStreamSynchronize(stream));
    }
    st.unique_keys = (uint64_t)U;
    st.sort_ms = tsort.stop();
    st.launches += 2;

    // ---- 3. incidence bitmap over the values two or more sequences hold ----
    Timer tpair(stream);
    Dev<uint8_t> d_shared((size_t)U);
    Dev<uint32_t> d_head((size_t)U), d_head_incl((size_t)U);
    mash_mark_kernel<<<blocks_of(U, 256), 256, 0, stream>>>(U, d_ukeys.p, seq_bits, d_diag.p, d_shared.p, d_head.p);
    PGMM_CUDA(cudaGetLastError());
    uint32_t V2 = 0;
    {
      size_t bytes = 0;
      PGMM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, d_head.p, d_head_incl.p, U, stream));
      Dev<uint8_t> tmp(bytes);
      PGMM_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, bytes, d_head.p, d_head_incl.p, U, stream));
      PGMM_CUDA(cudaMemcpyAsync(&V2, d_head_incl.p + (U - 1), 4, cudaMemcpyDeviceToHost, stream));
      PGMM_CUDA(cudaStreamSynchronize(stream));
    }
    st.shared_values = V2;
    st.launches += 2;
    if (V2 > 0 && n > 1) {
      const uint64_t words = ((uint64_t)V2 + 63) / 64;
      size_t fr = 0, tot = 0;
      PGMM_CUDA(cudaMemGetInfo(&fr, &tot));
      if ((double)words * 8.0 * n > (double)fr * 0.9) {
        fprintf(stderr, "[pgmm_b200] mash distance: the incidence bitmap (%d sequences x %u shared values) needs %.1f GB, %.1f GB free\n", n,
                V2, (double)words * 8.0 * n / 1e9, (double)fr / 1e9);
        return -5;
      }
      Dev<unsigned long long> d_bitmap((size_t)(words * (uint64_t)n));
      PGMM_CUDA(cudaMemsetAsync(d_bitmap.p, 0, (size_t)(words * (uint64_t)n) * 8, stream));
      mash_bitmap_kernel<<<blocks_of(U, 256), 256, 0, stream>>>(U, d_ukeys.p, seq_bits, d_shared.p, d_head_incl.p, words, d_bitmap.p);
      PGMM_CUDA(cudaGetLastError());
      // ---- 4. pairs ----
      const int nb = (n + 3) / 4;
      dim3 grid((unsigned)((nb + kPairWarps - 1) / kPairWarps), (unsigned)nb);
      if (grid.y > 65535) return -5;
      mash_pair_kernel<<<grid, 32 * kPairWarps, 0, stream>>>(n, words, d_bitmap.p, d_counts.p);
      PGMM_CUDA(cudaGetLastError());
      PGMM_CUDA(cudaStreamSynchronize(stream));
      st.launches += 2;
    }
    mash_diag_kernel<<<blocks_of(n, 256), 256, 0, stream>>>(n, d_diag.p, d_counts.p);
    PGMM_CUDA(cudaGetLastError());
    PGMM_CUDA(cudaMemcpyAsync(counts, d_counts.p, (size_t)n * n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    PGMM_CUDA(cudaStreamSynchronize(stream));
    st.pair_ms = tpair.stop();
    st.launches += 1;
  }
  if (stats) *stats = st;
  return 0;
}

}  // namespace mash
}  // namespace pgmm
