// Ragged prefill / extend attention, third tcgen05 variant: 64-key K/V tiles with DOUBLE-BUFFERED score
// tiles in TMEM, so that QK^T of tile j+1 no longer waits for the softmax of tile j.
//
// Same contract as attn_prefill_tc.cu (replaces store_kv + BatchPrefillWithPagedKVCacheWrapper.run,
// python/minisgl/attention/fi.py:150-165,185-188; causal, bottom-right aligned; KV append fused).
//
// Why: in attn_prefill_tc.cu a sub-tile's chain  S ready -> softmax -> P ready -> [PV(j), QK(j+1)] -> S ready
// is strictly serial; the ncu line profile of 4096-token prompts (profiles/r02_ncu_prefill_cfg4_lines.txt)
// shows the softmax warps waiting for S 32 % of the time while the tensor pipe is 34 % busy -- the MMA
// turn-around (issue latency + two 128-wide GEMMs + mbarrier wake-ups) sits on the softmax's critical path.
// Here a K/V tile has 64 keys, a score tile is 128 x 64 fp32 = 64 TMEM columns, and each of the two query
// sub-tiles owns TWO score buffers:
//
//   TMEM (512 columns): S_A0 0-63 | S_A1 64-127 | S_B0 128-191 | S_B1 192-255 | O_A 256-383 | O_B 384-511
//   (P_s,b = packed 16-bit pairs over the first 32 columns of S_s,b)
//
// The UMMA issuer is event driven (non-blocking mbarrier probes): QK^T(s, j) is issued as soon as K tile j has
// landed and score buffer j & 1 is free (PV(s, j-2) issued -- the tensor pipe executes in order), PV(s, j) as
// soon as P(s, j) and V tile j are there.  The softmax warps therefore find S(j+1) waiting when they finish
// tile j; they only ever wait for the tensor core at the start of a unit.
//
//   warp 0      K producer (+ the unit's Q sub-tiles), 4-stage ring of 16 KB tiles   } TMA boxes per page piece
//   warp 1      V producer, 4-stage ring; fused KV append by TMA store from the tiles } / gather4 rows
//   warp 2      UMMA issuer (one thread) + TMEM allocation
//   warps 4-11  softmax, ONE thread per query row (TMEM lane) over the tile's 64 columns: no cross-thread
//               exchange, no named barrier; warps 4-7 sub-tile A, 8-11 sub-tile B
//
// Work unit = (request, 128-row query tile, kv head, pair of query heads) as before: the two heads share every
// K/V tile.  Exact causal flops = 4*Hq*D*sum_r[q*cached + q(q+1)/2]; the finer tiles also waste less of the
// masked diagonal (128 x 64 instead of 128 x 128 blocks).
#include "b200attn.h"
#include "common.cuh"
#include "sm100.cuh"

#include <type_traits>

namespace b200 {

int get_tensor_map_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows, bool is_bf16);

namespace pv3 {

using namespace sm100;

constexpr int kD = 128;
constexpr int kBM = 128;                    // query rows per sub-tile
constexpr int kBN = 64;                     // keys per tile
constexpr int kStages = 4;                  // K ring depth = V ring depth
constexpr int kThreads = 384;               // 4 control warps + 8 softmax warps
constexpr int kQHalfBytes = kBM * 128;      // [128 rows x 64 dims] 128B-swizzled: 16 KB
constexpr int kQTileBytes = 2 * kQHalfBytes;
constexpr int kHalfBytes = kBN * 128;       // [64 keys x 64 dims]: 8 KB
constexpr int kTileBytes = 2 * kHalfBytes;  // 16 KB
constexpr int kMaxUnitsSmem = 64;
constexpr int kTmemCols = 512;
constexpr float kRescaleThreshold = 8.0f;   // log2 domain

struct Smem {
  static constexpr int q = 0;                                // 2 sub-tiles x 32 KB
  static constexpr int kring = 2 * kQTileBytes;              // kStages x 16 KB
  static constexpr int vring = kring + kStages * kTileBytes;
  static constexpr int bars = vring + kStages * kTileBytes;  // 32 mbarriers
  static constexpr int tmem_ptr = bars + 32 * 8;
  static constexpr int units = tmem_ptr + 16;
  static constexpr int total = units + kMaxUnitsSmem * 48;
};
static_assert(Smem::total + 1024 <= 232448, "prefill v3 shared memory exceeds the 227 KB per-CTA limit");
// kSFull / kPFull / kOFull: index = sub * 2 + (tile & 1)
enum Bar { kFullK = 0, kEmptyK = 4, kFullV = 8, kEmptyV = 12, kQFull = 16, kQEmpty = 17, kSFull = 18, kPFull = 22, kOFull = 26 };

template <typename T>
struct Params {
  const int32_t* slot_table;
  int64_t st_stride;
  const int32_t* seq_lens;
  const int32_t* cu_q;
  const int32_t* plan;  // {n_items, 0,0,0, item[...]}
  int bs, hq, hkv;
  int num_slots;
  int box_rows;
  float scale_log2;
  T* out;
  const T* k_new;
  const T* v_new;
  int64_t kv_rs;
  T* k_cache;
  T* v_cache;
  const int32_t* out_loc;
  int64_t nnz;
  int skip_append;
};

struct Unit {
  int r, q_start, q_len, q_begin, kv_len, cached, kv_hi, n_tiles, h, head0, n_sub;
};
static_assert(sizeof(Unit) <= 48, "Unit must fit its smem slot");

__device__ __forceinline__ Unit get_unit(int pos, int hkv, int group, const int32_t* items,
                                         const int32_t* seq_lens, const int32_t* cu_q) {
  Unit u;
  const int n_pairs = (group + 1) >> 1;
  const int per_item = hkv * n_pairs;
  const int item = pos / per_item;
  const int rem = pos - item * per_item;
  u.h = rem / n_pairs;
  const int pair = rem - u.h * n_pairs;
  const int e = items[item];
  u.r = e & 0xffff;
  u.q_start = ((e >> 16) & 0xffff) * kBM;
  u.q_begin = cu_q[u.r];
  u.q_len = cu_q[u.r + 1] - u.q_begin;
  u.kv_len = seq_lens[u.r];
  u.cached = u.kv_len - u.q_len;
  u.kv_hi = min(u.kv_len, u.cached + min(u.q_len, u.q_start + kBM));
  u.n_tiles = (u.kv_hi + kBN - 1) / kBN;
  u.head0 = u.h * group + pair * 2;
  u.n_sub = min(2, group - pair * 2);
  return u;
}

template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  typename DTypeTraits<T>::T2 v = DTypeTraits<T>::from_float2(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// tcgen05.ld of 32 columns into r[0..31] (a slice of a larger register array)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
attn_prefill_v3_kernel(const Params<T> p, const __grid_constant__ CUtensorMap map_q,
                       const __grid_constant__ CUtensorMap map_k,
                       const __grid_constant__ CUtensorMap map_v,
                       const __grid_constant__ CUtensorMap box_k,
                       const __grid_constant__ CUtensorMap box_v,
                       const __grid_constant__ CUtensorMap new_k,    // k input rows, box {64, 1}
                       const __grid_constant__ CUtensorMap new_v,
                       const __grid_constant__ CUtensorMap newbox_k, // k input rows, box {64, box_rows}
                       const __grid_constant__ CUtensorMap newbox_v) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid / kWarp, lane = tid % kWarp;
  auto bar = [&](int i) { return sbase + Smem::bars + i * 8; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + Smem::tmem_ptr);
  uint8_t* sUnits = smem + Smem::units;

  const int group = p.hq / p.hkv;
  const int n_pairs = (group + 1) >> 1;
  const int total_units = p.plan[0] * p.hkv * n_pairs;
  const int32_t* items = p.plan + 4;
  const int grid = gridDim.x, cta = blockIdx.x;
  const int n_rounds = (total_units + grid - 1) / grid;
  auto pos_of = [&](int k) { return k * grid + ((k & 1) ? grid - 1 - cta : cta); };

  // ---------------------------------------------------------------- one-time setup
  for (int k = tid; k < kMaxUnitsSmem && k < n_rounds; k += kThreads) {
    const int pos = pos_of(k);
    Unit u;
    u.n_tiles = -1;
    if (pos < total_units) u = get_unit(pos, p.hkv, group, items, p.seq_lens, p.cu_q);
    *reinterpret_cast<Unit*>(sUnits + k * 48) = u;
  }
  auto unit_at = [&](int k) {
    if (k < kMaxUnitsSmem) return *reinterpret_cast<const Unit*>(sUnits + k * 48);
    const int pos = pos_of(k);
    Unit u;
    u.n_tiles = -1;
    if (pos < total_units) u = get_unit(pos, p.hkv, group, items, p.seq_lens, p.cu_q);
    return u;
  };
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar(kFullK + s), 1);
      mbar_init(bar(kEmptyK + s), 1);
      mbar_init(bar(kFullV + s), 1);
      mbar_init(bar(kEmptyV + s), 1);
    }
    mbar_init(bar(kQFull), 1);
    mbar_init(bar(kQEmpty), 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(kSFull + i), 1);
      mbar_init(bar(kPFull + i), 4);  // one elected arrival per softmax warp of the sub-tile
      mbar_init(bar(kOFull + i), 1);  // PV(s, j) completes on barrier (s, j & 1)
    }
    fence_barrier_init();
    prefetch_tensormap(&map_q);
    prefetch_tensormap(&map_k);
    prefetch_tensormap(&map_v);
    prefetch_tensormap(&box_k);
    prefetch_tensormap(&box_v);
    prefetch_tensormap(&new_k);
    prefetch_tensormap(&new_v);
    prefetch_tensormap(&newbox_k);
    prefetch_tensormap(&newbox_v);
  }
  if (warp == 2) tmem_alloc(sbase + Smem::tmem_ptr, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp < 2) {
    // ============================================================ TMA producers: warp 0 = K (+Q), warp 1 = V
    const int kind = warp;
    const int rb = p.box_rows > kBN ? kBN : p.box_rows;  // rows per tiled box (<= 64), 0 = gather4 mode
    const CUtensorMap* gmap = kind == 0 ? &map_k : &map_v;
    const CUtensorMap* bmap = kind == 0 ? &box_k : &box_v;
    const int full0 = kind == 0 ? kFullK : kFullV, empty0 = kind == 0 ? kEmptyK : kEmptyV;
    const uint32_t ring = sbase + (kind == 0 ? Smem::kring : Smem::vring);
    uint32_t tile_count = 0, unit_count = 0;
    // ---- fused KV append: the tile of a unit that holds the unit's own new tokens is written to the pool
    // straight from shared memory (TMA store) -- no extra loads, no extra launch.  Ownership: positions
    // [cached + q_start, cached + min(q_len, q_start + 128)) of kv head h belong to the unit of query tile
    // q_start and the first head pair.
    struct Pending {
      int valid, lo, hi, tile_begin, col0;
      uint32_t phase;
      const int32_t* slots;
    } pend[kStages];
    for (int s = 0; s < kStages; ++s) pend[s].valid = 0;
    // valid: 0 = nothing, 1 = tile issued (append once it has landed), 2 = stores issued
    auto issue_stores = [&](int stage) {  // whole warp; the tile has landed
      Pending& pd = pend[stage];
      const uint32_t base = ring + stage * kTileBytes;
      if (rb > 0) {
        const int n_instr = (kBN / rb) * 2;
        if (lane < n_instr) {
          const int box = lane >> 1, half = lane & 1;
          const int pb = pd.tile_begin + box * rb;
          const uint32_t src = base + half * kHalfBytes + box * rb * 128;
          const int col = pd.col0 + half * 64;
          if (pb >= pd.lo && pb + rb <= pd.hi) {
            tma_store_2d(bmap, src, col, __ldg(pd.slots + pb));
          } else {
            for (int i = 0; i < rb; ++i)
              if (pb + i >= pd.lo && pb + i < pd.hi) tma_store_2d(gmap, src + i * 128, col, __ldg(pd.slots + pb + i));
          }
        }
      } else {
        const int grp = lane >> 1, half = lane & 1;  // 16 row groups x 2 halves
        const int pos0 = pd.tile_begin + grp * 4;
        int rr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          rr[i] = (pos0 + i >= pd.lo && pos0 + i < pd.hi) ? __ldg(pd.slots + pos0 + i) : p.num_slots;  // out of range: skipped
        if (pos0 + 4 > pd.lo && pos0 < pd.hi)
          tma_scatter4(gmap, base + half * kHalfBytes + grp * 512, pd.col0 + half * 64, rr[0], rr[1], rr[2], rr[3]);
      }
      bulk_commit();
      pd.valid = 2;
      __syncwarp();
    };
    auto try_store = [&](int stage) {
      if (pend[stage].valid == 1 && mbar_test_wait(bar(full0 + stage), pend[stage].phase)) issue_stores(stage);
    };
    auto flush_stage = [&](int stage) {  // before a ring slot is overwritten
      if (pend[stage].valid == 1) issue_stores(stage);
      if (pend[stage].valid == 2) {
        bulk_wait_read0();
        pend[stage].valid = 0;
      }
    };
    for (int k = 0; k < n_rounds; ++k) {
      const Unit u = unit_at(k);
      if (u.n_tiles <= 0) continue;
      const int32_t* slots = p.slot_table + (int64_t)u.r * p.st_stride;
      const int col0 = u.h * kD;
      const bool owner = !p.skip_append && (u.head0 == u.h * group);
      const int own_lo = u.cached + u.q_start, own_hi = u.cached + min(u.q_len, u.q_start + kBM);
      auto load_q = [&]() {
        // rows past the tensor end are zero filled, rows past the request's end belong to the next request:
        // computed, never stored
        mbar_wait(bar(kQEmpty), (unit_count & 1) ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(bar(kQFull), (uint32_t)u.n_sub * kQTileBytes);
          for (int s = 0; s < u.n_sub; ++s)
            for (int half = 0; half < 2; ++half)
              tma_load_2d(sbase + Smem::q + s * kQTileBytes + half * kQHalfBytes, &map_q, bar(kQFull),
                          (u.head0 + s) * kD + half * 64, u.q_begin + u.q_start);
        }
        __syncwarp();
      };
      // Positions < cached come from the pool through the slot table; positions >= cached are the tokens of this
      // very forward and are read straight from the k / v inputs (row q_begin + pos - cached), so the attention
      // never depends on the append performed concurrently.
      const CUtensorMap* nmap = kind == 0 ? &new_k : &new_v;
      const CUtensorMap* nbmap = kind == 0 ? &newbox_k : &newbox_v;
      auto load_rows_one_by_one = [&](uint32_t dst, uint32_t fb, int col, int pos0, int n) {
        for (int i = 0; i < n; ++i) {
          const int pos = pos0 + i;
          if (pos < u.cached) tma_load_2d(dst + i * 128, gmap, fb, col, __ldg(slots + pos));
          else tma_load_2d(dst + i * 128, nmap, fb, col, u.q_begin + pos - u.cached);
        }
      };
      for (int t = 0; t < u.n_tiles; ++t, ++tile_count) {
        const uint32_t stage = tile_count % kStages, phase = (tile_count / kStages) & 1;
        const int tile_begin = t * kBN;
        if (rb > 0) {
          const int n_instr = (kBN / rb) * 2;
          const int box = lane >> 1, half = lane & 1;
          const int pb = tile_begin + box * rb;
          const bool old_box = pb + rb <= u.cached, new_box = pb >= u.cached;
          int first_slot = p.num_slots;
          if (lane < n_instr && old_box) first_slot = __ldg(slots + pb);
          mbar_wait(bar(empty0 + stage), phase ^ 1);
          flush_stage(stage);
          if (lane == 0) mbar_arrive_expect_tx(bar(full0 + stage), kTileBytes);
          __syncwarp();
          if (lane < n_instr) {
            const uint32_t dst = ring + stage * kTileBytes + half * kHalfBytes + box * rb * 128;
            const uint32_t fb = bar(full0 + stage);
            const int col = col0 + half * 64;
            if (old_box) tma_load_2d(dst, bmap, fb, col, first_slot);
            else if (new_box) tma_load_2d(dst, nbmap, fb, col, u.q_begin + pb - u.cached);
            else load_rows_one_by_one(dst, fb, col, pb, rb);  // the box straddling cached_len
          }
        } else {
          // gather mode: 16 row groups x 2 halves = one instruction per lane
          mbar_wait(bar(empty0 + stage), phase ^ 1);
          flush_stage(stage);
          if (lane == 0) mbar_arrive_expect_tx(bar(full0 + stage), kTileBytes);
          __syncwarp();
          const int grp = lane >> 1, half = lane & 1;
          const int pos0 = tile_begin + grp * 4;
          const uint32_t dst = ring + stage * kTileBytes + half * kHalfBytes + grp * 512;
          const uint32_t fb = bar(full0 + stage);
          const int col = col0 + half * 64;
          if (pos0 + 4 <= u.cached) {
            const int4 rr = __ldg(reinterpret_cast<const int4*>(slots + pos0));
            tma_gather4(dst, gmap, fb, col, rr.x, rr.y, rr.z, rr.w);
          } else if (pos0 >= u.cached) {
            const int r0 = u.q_begin + pos0 - u.cached;
            tma_gather4(dst, nmap, fb, col, r0, r0 + 1, r0 + 2, r0 + 3);
          } else {
            load_rows_one_by_one(dst, fb, col, pos0, 4);
          }
        }
        if (kind == 0 && t == 0) load_q();  // after the first K tile: the Q buffer frees up later than a K slot
        for (int s2 = 0; s2 < kStages; ++s2)
          if (s2 != (int)stage) try_store(s2);
        // does this tile hold rows this unit has to append?
        const int lo = max(own_lo, tile_begin), hi = min(own_hi, tile_begin + kBN);
        if (owner && lo < hi) {
          pend[stage].valid = 1;
          pend[stage].lo = lo;
          pend[stage].hi = hi;
          pend[stage].tile_begin = tile_begin;
          pend[stage].col0 = col0;
          pend[stage].phase = phase;
          pend[stage].slots = slots;
        }
      }
      ++unit_count;
    }
    for (int s = 0; s < kStages; ++s)
      if (pend[s].valid == 1) {
        mbar_wait(bar(full0 + s), pend[s].phase);
        issue_stores(s);
      }
    bulk_wait0();
  } else if (warp == 2) {
    // ============================================================ UMMA issuer (one thread), event driven
    if (lane == 0) {
      constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
      constexpr uint32_t idesc_qk = make_idesc_f16(128, kBN, kBf16, false, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, 128, kBf16, false, true);  // B = V, MN-major
      uint32_t tile_base = 0, unit_count = 0;
      uint32_t p_uses[4] = {0, 0, 0, 0};  // P tiles consumed per (sub, buffer): phase of kPFull
      for (int k = 0; k < n_rounds; ++k) {
        const Unit u = unit_at(k);
        if (u.n_tiles <= 0) continue;
        mbar_wait(bar(kQFull), unit_count & 1);
        ++unit_count;
        tc_fence_after_sync();
        const int n = u.n_tiles, ns = u.n_sub;
        int qk_cnt[2] = {0, 0}, pv_cnt[2] = {0, 0};
        uint32_t spins = 0;
        while (pv_cnt[0] < n || (ns > 1 && pv_cnt[1] < n)) {
          bool progress = false;
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            if (s >= ns) continue;
            // ---- O_s (+)= P_s(j) . V(j)
            if (pv_cnt[s] < qk_cnt[s]) {
              const int j = pv_cnt[s], b = j & 1;
              const uint32_t gt = tile_base + j, stage = gt % kStages;
              if (mbar_test_wait(bar(kPFull + s * 2 + b), p_uses[s * 2 + b] & 1) &&
                  mbar_test_wait(bar(kFullV + stage), (gt / kStages) & 1)) {
                tc_fence_after_sync();
                ++p_uses[s * 2 + b];
                const uint32_t vb = sbase + Smem::vring + stage * kTileBytes;
                const uint32_t pa = tmem_base + s * 128 + b * 64;  // packed P over the first 32 columns
                const uint32_t d = tmem_base + 256 + s * 128;
#pragma unroll
                for (int kk = 0; kk < kBN / 16; ++kk) {
                  // B = V (MN-major): 16 keys = two 8-key swizzle atoms (1024 B each); dims 64..127 at +8 KB
                  const uint64_t db = make_smem_desc(vb + kk * 2048, kHalfBytes, 1024, kLayoutSW128);
                  umma_f16_ts(d, pa + kk * 8, db, idesc_pv, (j > 0) || kk > 0);
                }
                umma_commit(bar(kOFull + s * 2 + b));
                ++pv_cnt[s];
                // the V tile is dead once every sub-tile has issued its PV for it
                if (pv_cnt[0] > j && (ns == 1 || pv_cnt[1] > j)) umma_commit(bar(kEmptyV + stage));
                progress = true;
              }
            }
            // ---- S_s(j) = Q_s . K(j)^T   (score buffer j & 1 is free once PV(s, j-2) has been issued)
            if (qk_cnt[s] < n && qk_cnt[s] < pv_cnt[s] + 2) {
              const int j = qk_cnt[s], b = j & 1;
              const uint32_t gt = tile_base + j, stage = gt % kStages;
              if (mbar_test_wait(bar(kFullK + stage), (gt / kStages) & 1)) {
                tc_fence_after_sync();
                const uint32_t kb = sbase + Smem::kring + stage * kTileBytes;
                const uint32_t qa = sbase + Smem::q + s * kQTileBytes;
                const uint32_t d = tmem_base + s * 128 + b * 64;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                  const uint64_t da = make_smem_desc(qa + (kk >> 2) * kQHalfBytes + (kk & 3) * 32, 16, 1024, kLayoutSW128);
                  const uint64_t db = make_smem_desc(kb + (kk >> 2) * kHalfBytes + (kk & 3) * 32, 16, 1024, kLayoutSW128);
                  umma_f16_ss(d, da, db, idesc_qk, kk > 0);
                }
                umma_commit(bar(kSFull + s * 2 + b));
                ++qk_cnt[s];
                if (qk_cnt[0] > j && (ns == 1 || qk_cnt[1] > j)) {
                  umma_commit(bar(kEmptyK + stage));               // K tile j consumed by every sub-tile
                  if (j == n - 1) umma_commit(bar(kQEmpty));       // last QK of the unit
                }
                progress = true;
              }
            }
          }
          if (progress) spins = 0;
          else if (++spins > (1u << 28)) __trap();
        }
        tile_base += n;
      }
    }
  } else if (warp >= 4) {
    // ============================================================ softmax: 8 warps, one thread per query row.
    const int sub = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + lane;         // query row of the sub-tile = TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t s_base = tmem_base + lane_base + sub * 128;        // + buffer * 64
    const uint32_t o_addr = tmem_base + lane_base + 256 + sub * 128;  // my 128 output columns
    uint32_t s_uses[2] = {0, 0};  // score tiles consumed per buffer (phase of kSFull)
    // PV(j) completes on barrier kOFull[sub][j & 1].  o_deliv[b] = P tiles of parity b delivered so far (each is
    // followed by exactly one PV completion on barrier b), o_seen[b] = completions observed.  A barrier gets its
    // next completion only after the P two tiles later has been delivered, and PV(j-2) is observed before P(j) is
    // delivered, so at most one phase is ever pending -- the parity waits stay unambiguous -- while the wait
    // itself (a PV issued two softmax passes ago) never blocks.
    uint32_t o_deliv[2] = {0, 0}, o_seen[2] = {0, 0};
    auto observe = [&](int bb) {  // all PVs of parity bb whose P has been delivered are complete
      while (o_seen[bb] < o_deliv[bb]) {
        mbar_wait(bar(kOFull + sub * 2 + bb), o_seen[bb] & 1);
        ++o_seen[bb];
      }
    };
    for (int k = 0; k < n_rounds; ++k) {
      const Unit u = unit_at(k);
      if (u.n_tiles <= 0) continue;
      if (sub >= u.n_sub) continue;  // odd group size: the B warpgroup sits this unit out
      const int q_row = u.q_start + row;
      const int vis_end = min(u.cached + q_row + 1, u.kv_len);  // keys [0, vis_end) are visible
      float m_used = -INFINITY, l_run = 0.f;
      for (int j = 0; j < u.n_tiles; ++j) {
        const int b = j & 1, tile_begin = j * kBN;
        mbar_wait(bar(kSFull + sub * 2 + b), s_uses[b] & 1);
        ++s_uses[b];
        tc_fence_after_sync();
        const uint32_t s_addr = s_base + b * 64;
        // visible keys of this row inside the tile form a prefix [0, n_vis); rows of a warp are consecutive, so
        // lane 0 / lane 31 bound the warp: whole 32-column chunks are skipped, unmasked, or masked element-wise
        const int n_vis = max(0, min(kBN, vis_end - tile_begin));
        const int n_lo = __shfl_sync(0xffffffffu, n_vis, 0), n_hi = __shfl_sync(0xffffffffu, n_vis, 31);
        uint32_t s[64];
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (c * 32 < n_hi) tmem_ld32(s_addr + c * 32, s + c * 32);
        tmem_wait_ld();
        // ---- row max
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = c * 32;
          if (col0 >= n_hi) continue;
          if (col0 + 32 <= n_lo) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              mx0 = fmaxf(mx0, __uint_as_float(s[col0 + e]));
              mx1 = fmaxf(mx1, __uint_as_float(s[col0 + e + 1]));
              mx2 = fmaxf(mx2, __uint_as_float(s[col0 + e + 2]));
              mx3 = fmaxf(mx3, __uint_as_float(s[col0 + e + 3]));
            }
          } else {
            const int nv = n_vis - col0;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              mx0 = fmaxf(mx0, e < nv ? __uint_as_float(s[col0 + e]) : -INFINITY);
              mx1 = fmaxf(mx1, e + 1 < nv ? __uint_as_float(s[col0 + e + 1]) : -INFINITY);
              mx2 = fmaxf(mx2, e + 2 < nv ? __uint_as_float(s[col0 + e + 2]) : -INFINITY);
              mx3 = fmaxf(mx3, e + 3 < nv ? __uint_as_float(s[col0 + e + 3]) : -INFINITY);
            }
          }
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
        // ---- keep the kOFull phases in step: PV(j-2) (same parity; issued two passes ago, never blocks)
        observe(b);
        const bool grow = mx > m_used + kRescaleThreshold;   // also true for the first tile (-inf)
        const float m_new = grow ? mx : m_used;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          // rare (the running max moved by more than 2^8): O must be quiescent -- PV(j-1) complete, and PV(j)
          // cannot start before P(j) is delivered below
          observe(b ^ 1);
          tc_fence_after_sync();
          const float alpha = grow ? fast_exp2(m_used - m_new) : 1.f;
          l_run *= alpha;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_x32(o_addr + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st_x32(o_addr + c * 32, o);
          }
          tmem_wait_st();
        }
        m_used = m_new;
        const float m_sub = (m_used == -INFINITY) ? 0.f : m_used;
        // ---- P = exp2(S*scale - m), row sum, packed 16-bit pairs back to TMEM over the buffer's columns 0-31
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = c * 32;
          uint32_t pk[16];
          if (col0 >= n_hi) {
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = 0u;
          } else if (col0 + 32 <= n_lo) {
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const float p0 = fast_exp2(fmaf(__uint_as_float(s[col0 + e]), p.scale_log2, -m_sub));
              const float p1 = fast_exp2(fmaf(__uint_as_float(s[col0 + e + 1]), p.scale_log2, -m_sub));
              l0 += p0;
              l1 += p1;
              pk[e >> 1] = pack2<T>(p0, p1);
            }
          } else {
            const int nv = n_vis - col0;
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const float p0 = e < nv ? fast_exp2(fmaf(__uint_as_float(s[col0 + e]), p.scale_log2, -m_sub)) : 0.f;
              const float p1 = e + 1 < nv ? fast_exp2(fmaf(__uint_as_float(s[col0 + e + 1]), p.scale_log2, -m_sub)) : 0.f;
              l0 += p0;
              l1 += p1;
              pk[e >> 1] = pack2<T>(p0, p1);
            }
          }
          tmem_st_x16(s_addr + c * 16, pk);
        }
        l_run += l0 + l1;
        tmem_wait_st();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kPFull + sub * 2 + b));  // every lane's P is in TMEM
        ++o_deliv[b];
      }
      // ---- epilogue: O / l -> out (my row, 128 columns = 256 contiguous bytes); every PV of the unit complete
      observe(0);
      observe(1);
      tc_fence_after_sync();
      const float inv = 1.f / l_run;
      const bool store = q_row < u.q_len;
      T* orow = p.out + ((int64_t)(u.q_begin + q_row) * p.hq + (u.head0 + sub)) * kD;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t o[32];
        tmem_ld_x32(o_addr + c * 32, o);
        tmem_wait_ld();
        if (store) {
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            Vec8 w;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              w.w[e] = pack2<T>(__uint_as_float(o[v4 * 8 + 2 * e]) * inv, __uint_as_float(o[v4 * 8 + 2 * e + 1]) * inv);
            *reinterpret_cast<Vec8*>(orow + c * 32 + v4 * 8) = w;
          }
        }
      }
      // my O reads are complete (tmem_wait_ld); the next unit's PV(0), which overwrites O, is gated on kPFull,
      // which every warp of this warpgroup arrives on only after all its lanes finished this epilogue
      tc_fence_before_sync();
      __syncwarp();
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <typename T>
static int launch(const Params<T>& p, const void* q, int64_t q_rs, int64_t nnz, const void* k_cache,
                  const void* v_cache, cudaStream_t st) {
  CUtensorMap nk, nv, nbk, nbv;
  const bool bf16 = std::is_same<T, __nv_bfloat16>::value;
  CUtensorMap mq, mk, mv, bk, bv;
  const uint64_t cols = (uint64_t)p.hkv * kD;
  const int rb = p.box_rows > kBN ? kBN : p.box_rows;
  if (int rc = encode_tensor_map_2d(&mq, q, nnz, (uint64_t)p.hq * kD, q_rs * 2, 64, kBM, bf16)) return rc;
  if (int rc = get_tensor_map_2d(&mk, k_cache, p.num_slots, cols, cols * 2, 64, 1, bf16)) return rc;
  if (int rc = get_tensor_map_2d(&mv, v_cache, p.num_slots, cols, cols * 2, 64, 1, bf16)) return rc;
  bk = mk;
  bv = mv;
  if (rb > 0) {
    if (int rc = get_tensor_map_2d(&bk, k_cache, p.num_slots, cols, cols * 2, 64, rb, bf16)) return rc;
    if (int rc = get_tensor_map_2d(&bv, v_cache, p.num_slots, cols, cols * 2, 64, rb, bf16)) return rc;
  }
  // the k / v inputs of this forward as [nnz, hkv*128] row-strided tensors (activations move between forwards)
  if (int rc = encode_tensor_map_2d(&nk, p.k_new, nnz, cols, p.kv_rs * 2, 64, 1, bf16)) return rc;
  if (int rc = encode_tensor_map_2d(&nv, p.v_new, nnz, cols, p.kv_rs * 2, 64, 1, bf16)) return rc;
  nbk = nk;
  nbv = nv;
  if (rb > 0) {
    if (int rc = encode_tensor_map_2d(&nbk, p.k_new, nnz, cols, p.kv_rs * 2, 64, rb, bf16)) return rc;
    if (int rc = encode_tensor_map_2d(&nbv, p.v_new, nnz, cols, p.kv_rs * 2, 64, rb, bf16)) return rc;
  }
  const size_t smem = Smem::total + 1024;
  static bool configured = false;
  if (!configured) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_v3_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  attn_prefill_v3_kernel<T><<<num_sms(), kThreads, smem, st>>>(p, mq, mk, mv, bk, bv, nk, nv, nbk, nbv);
  B200_POST_LAUNCH();
  return 0;
}

}  // namespace pv3

extern std::atomic<int> g_prefill_skip_append;

// entry used by launch_prefill_tc (attn_prefill_tc.cu) when the variant option selects it
int launch_prefill_v3(const void* q, int64_t q_rs, int64_t nnz, const void* k, const void* v, int64_t kv_rs,
                      void* k_cache, void* v_cache, const int32_t* out_loc, const int32_t* slot_table, int64_t st_stride,
                      const int32_t* seq_lens, const int32_t* cu_q, const int32_t* prefill_plan, int bs, int hq, int hkv,
                      int64_t num_slots, int box_rows, float scale_log2, void* out, int dtype, cudaStream_t st) {
#define RUN(T_)                                                                                       \
  pv3::Params<T_> p{slot_table, st_stride, seq_lens, cu_q, prefill_plan, bs, hq, hkv, (int)num_slots,  \
                    box_rows, scale_log2, (T_*)out, (const T_*)k, (const T_*)v, kv_rs, (T_*)k_cache,   \
                    (T_*)v_cache, out_loc, nnz, g_prefill_skip_append.load()};                         \
  return pv3::launch<T_>(p, q, q_rs, nnz, k_cache, v_cache, st)
  if (dtype == B200_DTYPE_BF16) {
    RUN(__nv_bfloat16);
  }
  RUN(__half);
#undef RUN
}

}  // namespace b200
