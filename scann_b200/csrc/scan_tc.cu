// scan_tc.cu -- the LUT16 scan of DENSE batches on the tensor cores (tcgen05 kind::f8f6f4, TMEM accumulators).
//
// Replaces, for leaves that many queries of the batch probe (C2: 10k queries x 100 of 2000 leaves = 500 per leaf),
// the same reference code as scan.cu:
//   LUT16Avx2<>::GetTopFloatDistances / BottomLoop     hashes/internal/lut16_avx2.inc:55-124,404-527
//   FastTopNeighbors<float>                            utils/fast_top_neighbors.h:43-299
//
// The int16 sum of a (query, datapoint) pair is  sum_b lut[q][b][code[x][b]]  =  < lut[q][.], onehot(code[x][.]) >
// over K = 16 B positions: a GEMM of the queries' u8 LUTs with the one-hot expansion of the 4-bit codes.  The SIMT
// scan spends ~9 issue slots per eight lookups and tops out near 60 lookups / clk / SM.
//
// Number format.  kind::i8 is the natural fit, and it is exact, but on this part it runs ~12x below the 8-bit float
// rate (measured here: 1,600 clk per M128 N256 K32 instruction against 128 for e4m3; profiles/r02_scan_tc_notes.txt).
// So the LUT byte v is split into its nibbles, v = 16 h + l, and each nibble gets its own ROW of the A operand as an
// e4m3 number -- 16 h in {0, 16, .., 240} and l in {0, .., 15} are exactly representable (4 significant bits) -- while
// the one-hot operand holds e4m3 1.0.  Every product is an integer, every partial sum of a row is an integer below
// 2^11 x 16, so the f32 accumulation is exact whatever its order or internal width, and  sum = D[hi row] + D[lo row].
//
// Shape.  An item is (leaf, block of <= 64 queries).  A = 128 rows (64 queries x {hi, lo}; a query's two rows sit
// 16 lanes apart in the same TMEM lane quadrant) x K bytes, RESIDENT in shared memory for the whole leaf
// (SWIZZLE_128B K-major, W chunks of 16 KB, refilled chunk by chunk while the previous item's last tile drains).
// The leaf's datapoints stream through as B tiles of N <= 256 rows whose one-hot chunks (8 blocks = 128 bytes of K
// per stage) are generated straight from the packed code words by eight producer warps; one thread issues
// tcgen05.mma (M128 N K32, four per stage); the accumulator is double-buffered in TMEM (2 x 256 columns); four
// epilogue warps read it back with tcgen05.ld, add the two rows of a query with one shuffle per column, and append
// the survivors of `sum <= thr` / `key < tau` to the same per-query candidate buffers the SIMT scan fills.  The bits
// are identical by construction: same integer sums, same float conversion, same keys.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "exact_math.cuh"
#include "kernels.h"
#include "scan_common.cuh"

namespace sb {
namespace tc {

constexpr int TQ = 64;                // queries per item
constexpr int TM = 128;               // A rows = TMEM lanes: 64 queries x {hi, lo}
constexpr int TN = 256;               // datapoints per full tile = TMEM columns of one accumulator
constexpr int KC = 128;               // operand bytes of K per chunk: 8 blocks x 16 codes, one swizzle row
constexpr int kAChunk = TM * KC;      // 16 KB of the A operand per chunk
constexpr int kBStage = TN * KC;      // 32 KB per B stage
constexpr int kThreads = 512;         // warp 0 MMA, warp 2 TMEM alloc, warps 4-7 epilogue, warps 8-15 producers
constexpr int kProducerWarp0 = 8;
constexpr int kProducers = 256;       // one B row (datapoint) each; pairs of them build one A row each
constexpr int kStageKeys = 32;        // survivors an epilogue thread stages between flushes (a chunk adds at most 16)
constexpr int kTmemCols = 512;        // two 256-column f32 accumulators
constexpr int kMaxStages = 4;
constexpr int kMaxW = 9;
constexpr uint32_t kOneE4M3 = 0x38u;  // 1.0

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Waits carry a watchdog: a barrier that does not complete within ~2^24 polls (seconds; a stage takes microseconds)
// is a protocol bug -- the thread reports where it was stuck, raises a kernel-wide abort flag that every other wait
// sees, and the roles leave instead of hanging the GPU.
// The flag is word 7 of the batch's counters (zeroed per batch; the host reads the counters back and reports it).
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __noinline__ bool mbar_wait_slow(uint64_t* bar, uint32_t parity, volatile uint32_t* abort_flag, int code, uint32_t a,
                                            uint32_t b, uint32_t c) {
  for (uint32_t spins = 0;; ++spins) {
    if (mbar_try(bar, parity)) return true;
    if ((spins & 1023u) == 1023u && (spins > (1u << 24) || *abort_flag)) {
      if (!*abort_flag || (threadIdx.x & 31) == 0)
        printf("scan_tc watchdog: cta %d thread %d stuck at wait %d (item %u tile %u chunk %u, parity %u)\n", (int)blockIdx.x,
               (int)threadIdx.x, code, a, b, c, parity);
      *abort_flag = 1u;
      return false;
    }
  }
}
#define mbar_wait(bar, parity, code, a, b, c) (mbar_try(bar, parity) || mbar_wait_slow(bar, parity, w.counters + 7, code, a, b, c))
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// generic-proxy writes (st.shared) -> async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row atoms of 1024 B (SBO), LBO unused.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// kind::f8f6f4 instruction descriptor: D = f32 (1 << 4), A = B = e4m3 (0), both K-major, M x N.
__host__ __device__ constexpr uint32_t umma_idesc_e4m3(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- LUT bytes -> the two e4m3 operand rows of a query -------------------------------------------------------
// e4m3 of an integer n in 1..15: exponent e = floor(log2 n), byte = (e + 7) << 3 | top three bits below the leading
// one; of 16 n: exponent e + 4.  0 -> 0x00.
__host__ __device__ constexpr uint32_t e4m3_of_nibble(uint32_t n, uint32_t extra_exp) {
  if (n == 0) return 0u;
  uint32_t e = 0;
  while ((n >> (e + 1)) != 0) ++e;
  return ((e + 7u + extra_exp) << 3) | ((n << (3u - e)) & 7u);
}
__global__ void encode_lut_kernel(const uint8_t* __restrict__ lut, size_t nbytes, uint8_t* __restrict__ hi, uint8_t* __restrict__ lo) {
  __shared__ uint8_t t_hi[16], t_lo[16];
  if (threadIdx.x < 16) {
    t_hi[threadIdx.x] = (uint8_t)e4m3_of_nibble(threadIdx.x, 4u);
    t_lo[threadIdx.x] = (uint8_t)e4m3_of_nibble(threadIdx.x, 0u);
  }
  __syncthreads();
  const size_t nw = nbytes / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t v = reinterpret_cast<const uint32_t*>(lut)[i];
    uint32_t h = 0, l = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const uint32_t x = (v >> (8 * b)) & 255u;
      h |= (uint32_t)t_hi[x >> 4] << (8 * b);
      l |= (uint32_t)t_lo[x & 15u] << (8 * b);
    }
    reinterpret_cast<uint32_t*>(hi)[i] = h;
    reinterpret_cast<uint32_t*>(lo)[i] = l;
  }
}

// One work item: a leaf and a block of <= 64 of the queries that probe it.  The work list is built with 64 queries
// per item and whole leaves per item (scan_prepare_phase), so `item - item_off[leaf]` is the query block.
struct Item {
  uint32_t leaf, ebase, ecount, gbeg, ng, nleaf, ntiles;
};
__device__ __forceinline__ Item get_item(const DevIndex& ix, const ScanWork& w, uint32_t item) {
  Item it;
  if (w.item_leaf && item < w.item_leaf_cap) {
    it.leaf = w.item_leaf[item];
  } else {  // leaf = upper_bound(item_off, item) - 1
    uint32_t lo = 0, hi = ix.L;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (w.item_off[mid + 1] <= item) lo = mid + 1; else hi = mid;
    }
    it.leaf = lo;
  }
  const uint32_t chunk = item - w.item_off[it.leaf];
  it.ebase = w.leaf_eoff[it.leaf] + chunk * (uint32_t)TQ;
  it.ecount = min((uint32_t)TQ, w.leaf_eoff[it.leaf + 1] - it.ebase);
  it.gbeg = ix.leaf_goff[it.leaf];
  it.ng = ix.leaf_goff[it.leaf + 1] - it.gbeg;
  it.nleaf = ix.leaf_size[it.leaf];
  it.ntiles = (it.ng * 32u + (uint32_t)TN - 1u) / (uint32_t)TN;
  return it;
}
// rows of a tile (the leaf's last tile is ragged), in whole 32-slot groups: a multiple of the MMA's N granularity
__device__ __forceinline__ uint32_t tile_rows(const Item& I, uint32_t tile) {
  const uint32_t left = I.ng * 32u - tile * (uint32_t)TN;
  return left >= (uint32_t)TN ? (uint32_t)TN : left;
}

// one-hot stages that fit beside the resident LUT operand
__host__ __device__ constexpr int stages_for(int W) {
  return (216 * 1024 - W * kAChunk) / kBStage < kMaxStages ? (216 * 1024 - W * kAChunk) / kBStage : kMaxStages;
}

// 16 bytes of K with e4m3 1.0 at byte `nib`; nib8 = 8 * nib.  PTX shl clamps the shift amount at 32, so the words
// that do not hold the byte shift it out (amounts below zero wrap to huge unsigned values): no compares, no selects.
__device__ __forceinline__ uint32_t shl_one(uint32_t amount) {
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(kOneE4M3), "r"(amount));
  return r;
}
__device__ __forceinline__ uint4 onehot16(uint32_t nib8) {
  return make_uint4(shl_one(nib8), shl_one(nib8 - 32u), shl_one(nib8 - 64u), shl_one(nib8 - 96u));
}

template <int W>
__global__ void __launch_bounds__(kThreads, 1)
scan_tc_kernel(DevIndex ix, ScanWork w, const uint8_t* __restrict__ lut_hi, const uint8_t* __restrict__ lut_lo, int prof) {
  constexpr int n_stages = stages_for(W);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                          // [W][128 rows][128 B]
  uint8_t* sB = smem + (size_t)W * kAChunk;    // [n_stages][256 rows][128 B]
  __shared__ __align__(8) uint64_t a_full[kMaxW], a_empty[kMaxW], b_full[kMaxStages], b_empty[kMaxStages], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint32_t s_stage[kStageKeys][TM];  // per epilogue thread (column): (sum << 8 | column in tile) of survivors
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_items = w.counters[1];
  const int off128 = 128 * (int)ix.B;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kMaxW; ++i) { mbar_init(&a_full[i], kProducers); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&b_full[i], kProducers); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {  // ---- MMA issuer ----
      uint32_t bit = 0, lt = 0, ii = 0;
      long long t_a = 0, t_b = 0, t_t = 0, t0 = clock64(), c0;
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++ii) {
        const Item I = get_item(ix, w, item);
        for (uint32_t tile = 0; tile < I.ntiles; ++tile, ++lt) {
          const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
          const uint32_t idesc = umma_idesc_e4m3(TM, (int)tile_rows(I, tile));
          const bool last = tile + 1 == I.ntiles;
          c0 = clock64();
          if (!mbar_wait(&tmem_empty[as], aph ^ 1, 1, item, tile, lt)) return;  // the epilogue has drained this accumulator
          t_t += clock64() - c0;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tacc = tmem + as * TN;
#pragma unroll 1
          for (int kc = 0; kc < W; ++kc, ++bit) {
            const uint32_t st = bit % (uint32_t)n_stages, ph = (bit / (uint32_t)n_stages) & 1;
            if (tile == 0) {  // the item's LUT chunk kc has been written
              c0 = clock64();
              if (!mbar_wait(&a_full[kc], ii & 1, 2, item, tile, (uint32_t)kc)) return;
              t_a += clock64() - c0;
            }
            c0 = clock64();
            if (!mbar_wait(&b_full[st], ph, 3, item, tile, (uint32_t)kc)) return;
            t_b += clock64() - c0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t aaddr = smem_u32(sA + (size_t)kc * kAChunk);
            const uint32_t baddr = smem_u32(sB + (size_t)st * kBStage);
#pragma unroll
            for (int k = 0; k < KC / 32; ++k)
              umma_f8(tacc, umma_desc_sw128(aaddr + k * 32), umma_desc_sw128(baddr + k * 32), idesc, (kc | k) != 0 ? 1u : 0u);
            umma_commit(&b_empty[st]);           // frees the stage once these MMAs have read it
            if (last) umma_commit(&a_empty[kc]);  // ... and, on the item's last tile, the LUT chunk for the next item
          }
          umma_commit(&tmem_full[as]);
        }
      }
      if (prof && blockIdx.x == 0)
        printf("tc mma: items %u tiles %u total %lld clk, waits: a_full %lld b_full %lld tmem_empty %lld\n", ii, lt,
               clock64() - t0, t_a, t_b, t_t);
    }
  } else if (warp >= kProducerWarp0) {
    // ---- producers: the item's LUT operand (chunk by chunk, interleaved with the first tile), the one-hot tiles ----
    const uint32_t t = threadIdx.x - kProducerWarp0 * 32;  // 0..255: B row t; A row t & 127, K chunks of parity t >> 7
    const uint32_t sw = t & 7u;
    const uint32_t rowoff = (t >> 3) * 1024u + sw * 128u;
    // A row r: lane quadrant r >> 5 holds queries 16 (r >> 5) .. + 15, their hi rows in lanes 0-15, lo rows in 16-31
    const uint32_t ar = t & 127u, ahalf = t >> 7;
    const uint32_t aq = (ar >> 5) * 16u + (ar & 15u);
    const uint8_t* aplane = ((ar >> 4) & 1u) ? lut_lo : lut_hi;
    const uint32_t arow = smem_u32(sA) + (ar >> 3) * 1024u + (ar & 7u) * 128u;
    uint32_t bit = 0, ii = 0;
    long long t_a = 0, t_b = 0, t_gen = 0, t0 = clock64(), c0;
    // code words of (item, tile) for this thread's row; out of the leaf -> zeros
    auto fetch_codes = [&](const Item& I, uint32_t tile, uint32_t (&cw)[W]) {
      const uint32_t g = (tile * (uint32_t)TN + t) >> 5;
      if (g < I.ng) {
        load_codes<W>(ix.codes + (size_t)(I.gbeg + g) * W * 32, lane, cw);
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) cw[j] = 0u;
      }
    };
    Item I{};
    uint32_t item = blockIdx.x;
    uint32_t cw[W];
    if (item < n_items) { I = get_item(ix, w, item); fetch_codes(I, 0, cw); }
    for (; item < n_items; ++ii) {
      const bool has_q = aq < I.ecount;
      const uint4* asrc = reinterpret_cast<const uint4*>(aplane + (size_t)(has_q ? w.entry_q[I.ebase + aq] : 0u) * W * 128);
      // This thread copies the K chunks of its parity of LUT row `ar` with cp.async (16 bytes each, swizzled like the
      // TMA would): the first n_stages chunks up front, chunk kc + n_stages while one-hot chunk kc is generated, so a
      // chunk has n_stages stage times to land and the copies never wait for the previous item's MMAs.
      auto issue_a = [&](int kc) {
        if (kc >= W || (uint32_t)(kc & 1) != ahalf) return;
        c0 = clock64();
        if (!mbar_wait(&a_empty[kc], (ii & 1) ^ 1, 4, item, ii, (uint32_t)kc)) return;  // the previous item's last tile has read this chunk
        t_a += clock64() - c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t dstaddr = arow + (uint32_t)kc * kAChunk + (((uint32_t)j ^ (ar & 7u)) << 4);
          // src-size 0 zero-fills the 16 bytes (rows past the item's queries)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dstaddr), "l"(asrc + kc * 8 + j), "r"(has_q ? 16u : 0u) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
#pragma unroll
      for (int kc = 0; kc < n_stages; ++kc) issue_a(kc);
      const uint32_t next_item = item + gridDim.x;
      Item Inext = I;
      for (uint32_t tile = 0; tile < I.ntiles; ++tile) {
        const uint32_t nrows = tile_rows(I, tile);
        // the code words of the next tile (of this item or the next) are requested before this tile's chunks are generated
        uint32_t cn[W];
        if (tile + 1 < I.ntiles) {
          fetch_codes(I, tile + 1, cn);
        } else if (next_item < n_items) {
          Inext = get_item(ix, w, next_item);
          fetch_codes(Inext, 0, cn);
        }
#pragma unroll
        for (int kc = 0; kc < W; ++kc, ++bit) {
          if (tile == 0) issue_a(kc + n_stages);
          const uint32_t st = bit % (uint32_t)n_stages, ph = (bit / (uint32_t)n_stages) & 1;
          c0 = clock64();
          if (!mbar_wait(&b_empty[st], ph ^ 1, 5, item, tile, (uint32_t)kc)) return;
          t_b += clock64() - c0;
          c0 = clock64();
          if (t < nrows) {
            const uint32_t row = smem_u32(sB + (size_t)st * kBStage) + rowoff;
            const uint32_t word = cw[kc];
            // blocks past B (padding of the last code word) meet all-zero LUT rows: their one-hot bytes are harmless
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t nib8 = k == 0 ? (word << 3) & 0x78u : (word >> (4 * k - 3)) & 0x78u;
              sts128(row + (((uint32_t)k ^ sw) << 4), onehot16(nib8));
            }
          }
          if (tile == 0) {
            // LUT chunk kc has landed (groups complete in order; the groups of this thread's later chunks may be pending)
            if ((uint32_t)(kc & 1) == ahalf) {
              // own chunks committed after kc's group by now: kc + 2, kc + 4, .. up to kc + n_stages and W - 1
              const int later = (n_stages / 2) < ((W - 1 - kc) / 2) ? (n_stages / 2) : ((W - 1 - kc) / 2);
              if (later >= 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
              else if (later == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
              else asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            fence_proxy_async();
            mbar_arrive(&a_full[kc]);
          } else {
            fence_proxy_async();
          }
          mbar_arrive(&b_full[st]);
          t_gen += clock64() - c0;
        }
#pragma unroll
        for (int j = 0; j < W; ++j) cw[j] = cn[j];
      }
      I = Inext;
      item = next_item;
    }
    if (prof && blockIdx.x == 0 && t == 0)
      printf("tc producer: total %lld clk, waits: a_empty %lld b_empty %lld, one-hot work %lld\n",
             clock64() - t0, t_a, t_b, t_gen);
  } else if (warp >= 4) {
    // ---- epilogue: lanes l and l + 16 of a warp hold the hi and lo row of query 16 quad + l ----
    const int quad = warp & 3;
    const uint32_t row = (uint32_t)quad * 32u + (uint32_t)lane;  // TMEM lane = this thread's stage column
    const uint32_t qi = (uint32_t)quad * 16u + ((uint32_t)lane & 15u);
    const uint32_t half = (uint32_t)lane >> 4;                   // which 16 columns of a 32-column chunk this lane filters
    uint32_t lt = 0;
    long long t_w = 0, t0 = clock64(), c0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item I = get_item(ix, w, item);
      const bool qvalid = qi < I.ecount;
      uint32_t qq = 0;
      uint64_t tau = 0;
      float inv = 0.f, bias = 0.f;
      int thr = -0x7FFFFFFF;
      if (qvalid) {
        qq = w.entry_q[I.ebase + qi];
        bias = ix.key_by_dp ? 0.f : w.entry_bias[I.ebase + qi];
        tau = w.tau[qq];
        inv = w.inv_mult[qq];
        thr = acc_threshold(tau, w.mult[qq], inv, bias) + off128;
        if (prof >= 2) thr = -0x7FFFFFFF;  // timing experiments: no survivors
      }
      uint64_t* dst = w.buf + (size_t)qq * w.cap;
      // Sums that pass the integer pre-filter are staged as (sum << 17 | slot in leaf) in the thread's own
      // shared-memory column.  The stage is flushed by the whole warp at once -- exact key test, ONE reservation per
      // thread in its query's buffer, the atomics of all lanes in flight together -- when some lane's stage could
      // overflow in the next chunk, and at the end of the item (C2: ~4 survivors per query and item).
      const bool wide_leaf = I.ng * 32u > (1u << 17);  // slots do not fit the packed form: flush every tile, slot relative to it
      uint32_t ns = 0, flush_slot0 = 0;
      auto flush = [&]() {
        uint32_t n_ok = 0;
        for (uint32_t e = 0; e < ns; ++e) {
          const uint32_t pk = s_stage[e][row];
          const uint32_t col = flush_slot0 + (pk & 0x1FFFFu);
          bool ok = col < I.nleaf;  // columns past the leaf's real slots do not exist
          if (ok) {
            const uint32_t gslot = I.gbeg * 32u + col;
            const uint64_t key = make_key(ah_float_score((int)(pk >> 17) - off128, inv, bias), ix.key_by_dp ? ix.slot_dp[gslot] : gslot);
            ok = key < tau;
          }
          if (ok) s_stage[n_ok++][row] = pk;  // compact the survivors of the exact test in place
        }
        if (n_ok) {
          uint32_t pos = atomicAdd(&w.cnt[qq], n_ok);
          if (pos + n_ok > w.cap) w.ovf[qq] = 1u;
          for (uint32_t e = 0; e < n_ok; ++e, ++pos) {
            const uint32_t pk = s_stage[e][row];
            const uint32_t gslot = I.gbeg * 32u + flush_slot0 + (pk & 0x1FFFFu);
            if (pos < w.cap)
              dst[pos] = make_key(ah_float_score((int)(pk >> 17) - off128, inv, bias), ix.key_by_dp ? ix.slot_dp[gslot] : gslot);
          }
        }
        ns = 0;
      };
      for (uint32_t tile = 0; tile < I.ntiles; ++tile, ++lt) {
        const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
        const uint32_t nchunks = prof == 3 ? 0u : (tile_rows(I, tile) + 31u) / 32u;  // 3: the epilogue reads nothing
        c0 = clock64();
        if (!mbar_wait(&tmem_full[as], aph, 6, item, tile, lt)) return;
        t_w += clock64() - c0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tbase = tmem + as * TN + ((uint32_t)(quad * 32) << 16);
        const uint32_t slot0 = wide_leaf ? 0u : tile * (uint32_t)TN;
        if (wide_leaf) flush_slot0 = tile * (uint32_t)TN;
        // One pass over the tile in 32-column chunks (not unrolled: three roles share the instruction cache); one
        // shuffle per column pairs a query's hi and lo sums.
#pragma unroll 1
        for (uint32_t c = 0; c < nchunks; ++c) {
          uint32_t v[32];
          tmem_ld32(tbase + c * 32u, v);
          int s[16];
          uint32_t m = 0;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // lanes 0-15 keep columns 0-15 and hand their hi sums of columns 16-31 to the partner, and vice versa
            const uint32_t send = half ? v[j] : v[j + 16];
            const uint32_t recv = __shfl_xor_sync(0xFFFFFFFFu, send, 16);
            const uint32_t own = half ? v[j + 16] : v[j];
            s[j] = __float2int_rn(__fadd_rn(__uint_as_float(own), __uint_as_float(recv)));
            if (s[j] <= thr) m |= 1u << j;
          }
          if (m) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if ((m >> j) & 1u) {
                s_stage[ns][row] = ((uint32_t)s[j] << 17) | (slot0 + c * 32u + half * 16u + (uint32_t)j);
                ++ns;
              }
            }
          }
          __syncwarp();
          if (__any_sync(0xFFFFFFFFu, ns > (uint32_t)(kStageKeys - 16))) flush();  // the next chunk adds up to 16
        }
        // the accumulator is free: hand it back to the MMA warp
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[as]);
        if (wide_leaf && __any_sync(0xFFFFFFFFu, ns != 0)) flush();
      }
      __syncwarp();
      if (__any_sync(0xFFFFFFFFu, ns != 0)) flush();
    }
    if (prof && blockIdx.x == 0 && row == 0) printf("tc epilogue: total %lld clk, wait tmem_full %lld\n", clock64() - t0, t_w);
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
  }
}

template <int W>
static cudaError_t launch_t(const DevIndex& ix, const ScanWork& w, cudaStream_t s) {
  constexpr int stages = stages_for(W);
  const size_t smem = (size_t)W * kAChunk + (size_t)stages * kBStage + 1024;
  cudaError_t e = cudaFuncSetAttribute(scan_tc_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t lut_bytes = (size_t)w.nq * W * 128;
  encode_lut_kernel<<<sms * 4, 256, 0, s>>>(w.lut, lut_bytes, w.lut_e4m3, w.lut_e4m3 + lut_bytes);
  const char* pe = getenv("SCANN_B200_TC_PROFILE");
  scan_tc_kernel<W><<<sms, kThreads, smem, s>>>(ix, w, w.lut_e4m3, w.lut_e4m3 + lut_bytes, pe ? atoi(pe) : 0);
  return cudaGetLastError();
}

}  // namespace tc

// W <= 9: the resident LUT operand (W x 16 KB) and at least two one-hot stages fit one SM's shared memory
bool scan_tc_supported(const DevIndex& ix) { return ix.W >= 1 && ix.W <= tc::kMaxW && tc::stages_for((int)ix.W) >= 2; }
uint32_t scan_tc_queries_per_item() { return (uint32_t)tc::TQ; }

cudaError_t launch_scan_tc(const DevIndex& ix, const ScanWork& w, cudaStream_t s) {
  if (!w.lut_e4m3) return cudaErrorInvalidValue;
  switch (ix.W) {
    case 1: return tc::launch_t<1>(ix, w, s);
    case 2: return tc::launch_t<2>(ix, w, s);
    case 3: return tc::launch_t<3>(ix, w, s);
    case 4: return tc::launch_t<4>(ix, w, s);
    case 5: return tc::launch_t<5>(ix, w, s);
    case 6: return tc::launch_t<6>(ix, w, s);
    case 7: return tc::launch_t<7>(ix, w, s);
    case 8: return tc::launch_t<8>(ix, w, s);
    case 9: return tc::launch_t<9>(ix, w, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace sb
