// cco_rows2.cuh -- second-generation fused row kernel k_rows2 (H4 A'^T B' counts + H5 LLR + H6 top-k of SURVEY.md 8a;
// Mahout `drmA.t %*% drmB` + computeSimilarities, called from /root/reference/src/main/scala/URAlgorithm.scala:323-346).
//
// Same arithmetic, same results as k_rows (cco_kernels.cuh); what changed is how a row moves through the SM
// (profiles/r01_k_rows_ncu_full.md: warp bins ran at 14-39 % warps-active on 6.4-8.4 KB of fixed shared memory per
// warp, CTA bins lost 55 % of their stall samples at barriers, long_scoreboard 22 %):
//   * extents, not pointers: k_row_work leaves one (start, len) pair of B' per (item, user) in item-major order, so the
//     count phase reads ONE coalesced 8-byte stream per row (register-prefetched one chunk ahead) instead of the
//     at_users -> b_ptr[u], b_ptr[u+1] double gather; four B' column gathers in flight per lane.
//   * shared-memory diet: the level-1 cut histogram aliases the radix-select histogram (512 u16 bins), x11 table once
//     per CTA, 256-byte control block: 3.8 KB fixed per warp instead of 8.4 KB.
//   * compaction and the level-1 cut histogram are one pass over the table.
//   * the filter/evaluate loop has no per-round CTA barrier: warps reserve candidate slots with one atomic and keep
//     what does not fit pending; one barrier per row in the common case, and every decision that guards a barrier is
//     taken from a value nobody can write between two barriers (fixes the latent race of k_rows' `n = vctrl[0]`).
//   * no cliffs: a row whose counts do not fit the packed word, or whose hash table overflows, is split into more
//     residue passes (key' = b / n_pass, pass = b % n_pass) instead of failing the train.
#pragma once

#include "cco_kernels.cuh"

namespace cco {

constexpr int kCut2Bins = 512;       // level-1 cut: u16 bins of colB (aliases the 1 KB radix-select histogram)
constexpr int kCtrl2Ints = 64;       // [0] ncand [1] have_thr [2] hash overflow [4..7] threshold entry [9] cut1 [10] chunk cursor
                                     // [24..27] select state [40..55] dominance frontier
constexpr int kCtaShared2 = 256;     // x11tab (32 doubles), once per CTA
constexpr int kGather = 2;           // B' column gathers in flight per lane in the count phase

template <bool DENSE>
__device__ __forceinline__ void accumulate2(uint32_t *table, uint32_t tsize, uint32_t key, int cbits, int *ovf) {
  if (DENSE) {
    atomicAdd(&table[key], 1u);
    return;
  }
  uint32_t slot = __umulhi(key * 0x9e3779b1u, tsize);
  const uint32_t want = key << cbits;
  uint32_t probes = 0;
  while (true) {
    const uint32_t w = *reinterpret_cast<volatile uint32_t *>(&table[slot]);
    if ((w >> cbits) == key && w != kEmpty) { atomicAdd(&table[slot], 1u); return; }
    if (w == kEmpty) {
      const uint32_t old = atomicCAS(&table[slot], kEmpty, want | 1u);
      if (old == kEmpty) return;
      if ((old >> cbits) == key) { atomicAdd(&table[slot], 1u); return; }
    }
    slot = (slot + 1 == tsize) ? 0 : slot + 1;
    if (++probes > tsize) { *reinterpret_cast<volatile int *>(ovf) = 1; return; }
  }
}

// Final ordering of <= 64 candidates by one warp, in registers: lane l holds elements l and l + 32 of a 64-element
// bitonic network (best first, pads = key 0 sort last); partners at distance < 32 are exchanged with shuffles, the
// distance-32 step is lane-local.  r02 profile: the shared-memory bitonic sort was 30 % of the instructions of the
// smallest-row bin (two LDS.128 + two STS.128 + a warp barrier per compare-exchange).
__device__ __forceinline__ uint4 shfl_xor4(const uint4 &v, int j) {
  return make_uint4(__shfl_xor_sync(0xffffffffu, v.x, j), __shfl_xor_sync(0xffffffffu, v.y, j),
                    __shfl_xor_sync(0xffffffffu, v.z, j), __shfl_xor_sync(0xffffffffu, v.w, j));
}
__device__ __forceinline__ void sort_candidates_warp64(uint4 *tk, int n, int lane) {
  const uint4 pad = make_uint4(0u, 0u, 0xffffffffu, 0u);
  uint4 A = lane < n ? tk[lane] : pad, B = lane + 32 < n ? tk[lane + 32] : pad;
#pragma unroll
  for (int k2 = 2; k2 <= 64; k2 <<= 1) {
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      if (j == 32) {   // k2 == 64: (l, l + 32) live in the same lane, direction "up" everywhere
        if (cand_better(B, A)) { const uint4 t = A; A = B; B = t; }
      } else {
        const bool lower = (lane & j) == 0;
        const bool upA = k2 == 64 ? true : ((lane & k2) == 0);
        const bool upB = k2 == 64 ? true : (((lane + 32) & k2) == 0);
        const uint4 oA = shfl_xor4(A, j), oB = shfl_xor4(B, j);
        // the lower index of a pair takes the better element when the direction is "up", the worse one otherwise.
        // Keys are unique (columns are) except among pads, which are interchangeable: "not better" = "worse".
        const bool ob_a = cand_better(oA, A), ob_b = cand_better(oB, B);
        if ((lower == upA) == ob_a) A = oA;
        if ((lower == upB) == ob_b) B = oB;
      }
    }
  }
  __syncwarp();
  if (lane < n) tk[lane] = A;
  if (lane + 32 < n) tk[lane + 32] = B;
  __syncwarp();
}

template <int GROUP>
__device__ __forceinline__ bool group_and(bool p) {
  if (GROUP == 32) { __syncwarp(); return p; }   // warp-owned rows: `p` is warp-uniform by construction
  return __syncthreads_and(p ? 1 : 0) != 0;
}

template <int GROUP, bool DENSE>
__global__ void __launch_bounds__(GROUP == 32 ? 64 : GROUP, GROUP == 32 ? 16 : 1024 / GROUP) k_rows2(const RowArgs a) {
  const int GROUPS = GROUP == 32 ? (int)(blockDim.x >> 5) : 1;  // warp-owned rows: several independent warps per CTA
  constexpr int NW = GROUP / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int gid = tid / GROUP, gtid = tid % GROUP, gw = gtid >> 5;
  double *x11tab = reinterpret_cast<double *>(smem_raw);       // xLogX(j), j < 32: row independent, one per CTA
  unsigned char *base = smem_raw + kCtaShared2 + (size_t)gid * a.group_smem_bytes;
  uint4 *tk = reinterpret_cast<uint4 *>(base);                  // candidate buffer, a.cbuf entries
  uint4 *aux = tk + a.cbuf;                                      // a.caux entries (0 for warp-owned rows)
  double *x12tab = reinterpret_cast<double *>(aux + a.caux);    // 32 doubles
  int *ctrl = reinterpret_cast<int *>(x12tab + 32);             // kCtrl2Ints
  int *hist = ctrl + kCtrl2Ints;                                 // 256 ints: radix-select histogram / level-1 cut bins
  uint32_t *h1 = reinterpret_cast<uint32_t *>(hist);
  uint32_t *wqueue = reinterpret_cast<uint32_t *>(hist + 256);  // NW * 64 queued cells awaiting evaluation
  uint32_t *table = wqueue + NW * 64;
  volatile int *vctrl = ctrl;

  const int row_begin = a.bin_bounds[a.bin], row_end = a.bin_bounds[a.bin + 1];
  const bool varargs = (a.flags & CCO_FLAG_ENTROPY_VARARGS) != 0;
  const long long N = a.n_users;
  const double xN = xlogx(N);
  unsigned long long distinct_local = 0, evaluated_local = 0;
  if (tid < 32) x11tab[tid] = xlogx((long long)tid);
  __syncthreads();

  // The header of a row is a chain of dependent global loads (rows_sorted -> at_ptr / marg_a / row_work -> extents).  A warp
  // that owns a short row would expose all of it, so the next row's header is fetched while the current row is
  // processed: the item id at the top of the row, its pointers after the count phase, and (warp-owned rows) the first
  // chunk of extents before the final sort.
  const int rstride = gridDim.x * GROUPS;
  int ri = row_begin + blockIdx.x * GROUPS + gid;
  int item_n = -1, ra_n = 0;
  uint32_t ub_n = 0, ue_n = 0, w_n = 0;
  uint2 first_n = make_uint2(0u, 0u);
  if (ri < row_end) {
    item_n = a.rows_sorted[ri];
    ub_n = a.at_ptr[item_n];
    ue_n = a.at_ptr[item_n + 1];
    ra_n = a.marg_a[item_n];
    w_n = a.row_work[item_n];
    if (NW == 1 && !(a.tune & 3u) && ub_n + lane < ue_n) first_n = a.ext[ub_n + lane];
  }
  for (; ri < row_end; ri += rstride) {
    if (a.tune & 1u) {   // development switch: plain header loads at the top of the row
      item_n = a.rows_sorted[ri];
      ub_n = a.at_ptr[item_n];
      ue_n = a.at_ptr[item_n + 1];
      ra_n = a.marg_a[item_n];
      w_n = a.row_work[item_n];
    }
    const int item = item_n;
    const uint32_t u_begin = ub_n, u_end = ue_n;
    const long long ra = ra_n;
    const uint32_t w_row = w_n;
    const uint2 first = first_n;
    const bool pf_header = !(a.tune & 1u), pf_first = pf_header && !(a.tune & 2u);
    int item_nn = -1;
    if (ri + rstride < row_end) item_nn = a.rows_sorted[ri + rstride];   // consumed after the count phase
    bool header_fetched = false;
    bool first_ok = NW == 1 && pf_first;   // `first` holds this row's first extent chunk (used once, by the first count pass)
    const uint32_t dbound = w_row < (uint32_t)a.n_cols_b ? w_row : (uint32_t)a.n_cols_b;
    // packed word = (key' << cbits) | count with count <= min(rowA, max colB).  A row whose counts need more bits than
    // the launch-wide split leaves is cut into 2^extra residue passes: key' = b / n_pass loses `extra` bits.
    int cbits = a.count_bits;
    uint32_t n_pass = 1;
    {
      const uint32_t kmax = min((uint32_t)ra, (uint32_t)a.max_marg_b);
      if ((kmax + 1u) >> cbits) {
        const int extra = (32 - __clz(kmax + 1u)) - cbits;
        cbits += extra;
        n_pass = 1u << extra;
      }
    }
    if (!DENSE) n_pass = max(n_pass, (dbound + (uint32_t)a.cap - 1u) / (uint32_t)a.cap);
    if (n_pass == 0) n_pass = 1;
    unsigned long long distinct_row = 0;
    int emitted = 0;   // emit_all only: cells written so far (identical in every thread of the group)
    bool row_done = false;
    while (!row_done) {   // normally one trip; a hash-table overflow doubles n_pass and starts the row over
      row_done = true;
      const uint32_t cmask = (1u << cbits) - 1u;
      const uint32_t cols_per_pass = ((uint32_t)a.n_cols_b + n_pass - 1u) / n_pass;
      uint32_t tsize;
      if (DENSE) {
        tsize = cols_per_pass;
      } else {
        const uint32_t pb = dbound < cols_per_pass ? dbound : cols_per_pass;   // distinct keys one pass can see
        tsize = (pb > (uint32_t)a.cap) ? (uint32_t)a.slots
                                       : (uint32_t)min((unsigned long long)a.slots,
                                                       max(((unsigned long long)pb * (unsigned long long)a.tsize_x16) >> 4, 64ull));
        tsize = min((uint32_t)a.slots, (tsize + 32u * NW - 1u) / (32u * NW) * (32u * NW));
      }
      const bool use_cut = n_pass == 1 && w_row < 65536u && !a.emit_all;   // u16 bins cannot overflow; needs all cells at once
      group_sync<GROUP>();  // previous row (or attempt) fully done with shared memory
      if (gtid < 32) {
        const long long v = (gtid < kX12N) ? ra - gtid : N - ra;
        x12tab[gtid] = v >= 0 ? xlogx(v) : 0.0;
      }
      if (gtid == 0) { ctrl[0] = 0; ctrl[1] = 0; ctrl[2] = 0; ctrl[9] = 0x7fffffff; }
      if (gtid < 16) ctrl[40 + gtid] = 0x7fffffff;
      emitted = 0;
      distinct_row = 0;

      for (uint32_t pass = 0; pass < n_pass && row_done; ++pass) {
        // ---- clear ---------------------------------------------------------------------------------------
        for (uint32_t i = gtid; i < tsize; i += GROUP) table[i] = DENSE ? 0u : kEmpty;
        if (gtid == 0) ctrl[10] = 0;   // chunk cursor of the count phase
        if (use_cut)
          for (int i = gtid; i < kCut2Bins / 2; i += GROUP) h1[i] = 0u;
        group_sync<GROUP>();
        // ---- count: each warp takes 32-user chunks of the row's (start, len) extents; the products of a chunk are
        // flattened over the lanes; the next chunk's extents are already in flight -------------------------------
        {
          // warp-owned rows walk their users in order; in a CTA-owned row the warps draw chunks from a shared cursor
          // (ctrl[10]) so that a warp that met a heavy user (a long B' row) takes fewer chunks: the count phase ends
          // at a CTA barrier and r01 lost a quarter of its samples waiting there
          const uint32_t deg = u_end - u_begin;
          // CTA-owned rows: ~4 draws per warp (chunks of 8..32 users) keep the warps level at the barrier that ends the
          // count phase; 32-user chunks left 13 of 16 warps idle for half of it (r02 profile: 44 % of the samples)
          const uint32_t per = NW == 1 ? 32u : ((a.tune & 16u) ? min(32u, max(1u, (deg + NW - 1) / NW))
                                                               : min(32u, max(8u, (deg + 4 * NW - 1) / (4 * NW))));
          uint32_t c0;
          if (NW == 1) {
            c0 = u_begin;
          } else {
            uint32_t g = 0;
            if (lane == 0) g = atomicAdd(reinterpret_cast<uint32_t *>(&ctrl[10]), per);
            c0 = u_begin + __shfl_sync(0xffffffffu, g, 0);
          }
          uint2 nxt = make_uint2(0u, 0u);
          if (NW == 1 && first_ok) nxt = first;
          else if (lane < per && c0 + lane < u_end) nxt = a.ext[c0 + lane];
          first_ok = false;
          while (c0 < u_end) {
            const uint2 cur = nxt;
            uint32_t c1;
            if (NW == 1) {
              c1 = c0 + 32u;
            } else {
              uint32_t g = 0;
              if (lane == 0) g = atomicAdd(reinterpret_cast<uint32_t *>(&ctrl[10]), per);
              c1 = u_begin + __shfl_sync(0xffffffffu, g, 0);
            }
            nxt = make_uint2(0u, 0u);
            if (c1 < u_end && lane < per && c1 + lane < u_end) nxt = a.ext[c1 + lane];
            c0 = c1;
            const uint32_t s = cur.x, len = cur.y;
            uint32_t off = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const uint32_t v = __shfl_up_sync(0xffffffffu, off, d);
              if (lane >= d) off += v;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, off, 31);
            off -= len;  // exclusive
            for (uint32_t p0 = 0; p0 < total; p0 += 32 * kGather) {
              uint32_t bb[kGather];
              bool act[kGather];
#pragma unroll
              for (int h = 0; h < kGather; ++h) {
                const uint32_t p = p0 + h * 32 + lane;
                int j = 0;
#pragma unroll
                for (int st = 16; st > 0; st >>= 1) {
                  const int c = j + st;
                  const uint32_t v = __shfl_sync(0xffffffffu, off, c);
                  if (v <= p) j = c;
                }
                const uint32_t sj = __shfl_sync(0xffffffffu, s, j), oj = __shfl_sync(0xffffffffu, off, j);
                act[h] = p < total;
                bb[h] = act[h] ? (uint32_t)a.b_col[sj + (p - oj)] : 0u;
              }
#pragma unroll
              for (int h = 0; h < kGather; ++h) {
                if (!act[h]) continue;
                uint32_t key = bb[h];
                if (n_pass > 1) {
                  if (key % n_pass != pass) continue;
                  key /= n_pass;
                }
                accumulate2<DENSE>(table, tsize, key, cbits, &ctrl[2]);
              }
            }
          }
        }
        group_sync<GROUP>();
        if (!header_fetched && pf_header) {   // next row's pointers: item_nn has long arrived, these loads fly during compact/score
          header_fetched = true;
          item_n = item_nn;
          if (item_nn >= 0) {
            ub_n = a.at_ptr[item_nn];
            ue_n = a.at_ptr[item_nn + 1];
            ra_n = a.marg_a[item_nn];
            w_n = a.row_work[item_nn];
          }
        }
        if (!DENSE && vctrl[2]) {   // group-uniform: read between two barriers with no writer
          row_done = false;          // a pass saw more distinct keys than its table holds: split finer, start over
          break;
        }
        // ---- compact: each warp packs the occupied words of its own table segment, in place; the same pass feeds the
        // level-1 cut histogram: colB of the strongly positive k11 == 1 cells ---------------------------------------
        const uint32_t seg = (((tsize + NW - 1) / NW) + 31u) & ~31u;
        const uint32_t seg_lo = min((uint32_t)gw * seg, tsize), seg_hi = min(seg_lo + seg, tsize);
        uint32_t n_mine = 0;
        for (uint32_t pos = seg_lo; pos < seg_hi; pos += 32) {
          const uint32_t idx = pos + lane;
          uint32_t w = DENSE ? 0u : kEmpty;
          if (idx < seg_hi) w = table[idx];
          const bool valid = DENSE ? (w != 0u) : (w != kEmpty);
          const uint32_t word = DENSE ? ((idx << cbits) | w) : w;
          const unsigned m = __ballot_sync(0xffffffffu, valid);
          __syncwarp();
          if (valid) table[seg_lo + n_mine + __popc(m & ((1u << lane) - 1u))] = word;
          n_mine += __popc(m);
          __syncwarp();
        }
        // level-1 cut bins: colB of the strongly positive k11 == 1 cells; four gathers in flight per lane (one dependent
        // gather per trip left the warp waiting a full L2 round trip for every 32 cells)
        if (use_cut) {
          for (uint32_t q0 = 0; q0 < n_mine; q0 += 128) {
            uint32_t cbv[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const uint32_t q = q0 + h * 32 + lane;
              cbv[h] = 0xffffffffu;
              if (q < n_mine) {
                const uint32_t word = table[seg_lo + q];
                const uint32_t b = word >> cbits;   // n_pass == 1 here
                if ((word & cmask) == 1u && !(a.self && (int)b == item)) cbv[h] = (uint32_t)a.marg_b[b];
              }
            }
#pragma unroll
            for (int h = 0; h < 4; ++h)
              if (cbv[h] < (uint32_t)kCut2Bins && 2ull * (unsigned long long)ra * cbv[h] < (unsigned long long)N)
                atomicAdd(&h1[cbv[h] >> 1], 1u << (16u * (cbv[h] & 1u)));
          }
        }
        if (lane == 0) distinct_row += n_mine;
        if (a.emit_all) {
          // debug: every non-zero cell of the row (col, count), unordered
          int basepos = 0;
          if (lane == 0) basepos = atomicAdd(&ctrl[0], (int)n_mine);
          basepos = __shfl_sync(0xffffffffu, basepos, 0);
          for (uint32_t q = lane; q < n_mine; q += 32) {
            const uint32_t word = table[seg_lo + q];
            const size_t o = (size_t)item * a.out_stride + emitted + basepos + q;
            const uint32_t key = word >> cbits;
            a.out_col[o] = (int32_t)(n_pass > 1 ? key * n_pass + pass : key);
            a.out_cnt[o] = (int32_t)(word & cmask);
          }
          group_sync<GROUP>();
          emitted += vctrl[0];
          group_sync<GROUP>();
          if (gtid == 0) ctrl[0] = 0;
          group_sync<GROUP>();
          continue;
        }
        // ---- level-1 integer cut (exact): on the strongly positive side (2 rowA colB < N) the LLR of the k11 == 1
        // cells is strictly decreasing in colB, so once top_k of them sit at or below colB = c1, no k11 == 1 strongly
        // positive cell with colB > c1 can be kept: it is dropped by an integer compare in the filter stage.
        int cut1 = 0x7fffffff;
        if (use_cut) {
          group_sync<GROUP>();
          if (gtid < 32) {
            // lane l owns bins [16 l, 16 l + 16): 8 words
            uint32_t sum = 0;
#pragma unroll
            for (int wi = 0; wi < 8; ++wi) { const uint32_t v = h1[gtid * 8 + wi]; sum += (v & 0xffffu) + (v >> 16); }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
              if (gtid >= d) incl += v;
            }
            const uint32_t excl = incl - sum;
            int found = 0x7fffffff;
            if (excl < (uint32_t)a.top_k && incl >= (uint32_t)a.top_k) {
              uint32_t run = excl;
              for (int wi = 0; wi < 8; ++wi) {
                const uint32_t v = h1[gtid * 8 + wi];
                run += v & 0xffffu;
                if (run >= (uint32_t)a.top_k) { found = gtid * 16 + 2 * wi; break; }
                run += v >> 16;
                if (run >= (uint32_t)a.top_k) { found = gtid * 16 + 2 * wi + 1; break; }
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, o));
            if (gtid == 0) ctrl[9] = found;
          }
          group_sync<GROUP>();
          cut1 = vctrl[9];
        }
        // ---- score + select -----------------------------------------------------------------------------------
        const double x_ra = x12tab[0], x_nra = x12tab[kX12N];
        const double row_e = varargs ? __dsub_rn(xN, __dadd_rn(__dadd_rn(0.0, x_ra), x_nra))
                                     : __dsub_rn(__dsub_rn(xN, x_ra), x_nra);
        // Per warp: filter (integer only: diagonal, dominance frontier, level-1 cut) -> 64-entry queue -> evaluate 32
        // queued cells at a time in fp64 -> reserve candidate slots with one atomic.  What does not fit the buffer stays
        // pending in registers; the group then meets at ONE barrier, prunes, and resumes.
        uint32_t *wq = wqueue + gw * 64;
        uint32_t pos = 0;                            // next batch of 32 cells to LOAD
        int qn = 0;
        bool pend_ok = false;                       // this lane holds a candidate that found no slot yet
        uint4 e = make_uint4(0u, 0u, 0u, 0u);
        // one batch of cells (word, colB) is always in flight ahead of the filter
        uint32_t fword = 0;
        int fcb = 0;
        bool fvalid = false, fhave = false;
        if (n_mine > 0) {
          fvalid = lane < n_mine;
          if (fvalid) {
            fword = table[seg_lo + lane];
            const uint32_t key = fword >> cbits;
            fcb = a.marg_b[n_pass > 1 ? key * n_pass + pass : key];
          }
          pos = 32;
          fhave = true;
        }
        while (true) {
          bool done = false;
          while (true) {
            const unsigned pm = __ballot_sync(0xffffffffu, pend_ok);
            if (pm) {
              int basepos = 0;
              if (lane == 0) basepos = atomicAdd(&ctrl[0], __popc(pm));
              basepos = __shfl_sync(0xffffffffu, basepos, 0);
              const int idx = basepos + __popc(pm & ((1u << lane) - 1u));
              if (pend_ok && idx < a.cbuf) { tk[idx] = e; pend_ok = false; }
              if (__any_sync(0xffffffffu, pend_ok)) break;   // buffer full: wait for the prune
            }
            while (qn < 32 && fhave) {
              const uint32_t word = fword;
              const long long cb = fcb;
              const bool cv = fvalid;
              if (pos < n_mine) {
                const uint32_t q = pos + lane;
                fvalid = q < n_mine;
                if (fvalid) {
                  fword = table[seg_lo + q];
                  const uint32_t key = fword >> cbits;
                  fcb = a.marg_b[n_pass > 1 ? key * n_pass + pass : key];
                }
                pos += 32;
              } else {
                fhave = false;
              }
              bool surv = false;
              if (cv) {
                const uint32_t key = word >> cbits, k11 = word & cmask;
                const uint32_t b = n_pass > 1 ? key * n_pass + pass : key;
                if (!(a.self && (int)b == item)) {
                  // Dominance filter (exact): for fixed rowA and N, on the positively associated side (rowA*cb < k11*N) the LLR
                  // grows with k11 and shrinks with cb, so every evaluated cell (k, c) that fails strictly on LLR proves that
                  // all cells (k' <= k, c' >= c) fail too; cfail[k'] = smallest such c.
                  const bool pos_side = (unsigned long long)ra * (unsigned long long)cb < (unsigned long long)k11 * (unsigned long long)N;
                  surv = !(pos_side && k11 <= (uint32_t)kDomLevels && (int)cb >= vctrl[40 + k11]);
                  if (k11 == 1u && (int)cb > cut1 && 2ull * (unsigned long long)ra * (unsigned long long)cb < (unsigned long long)N)
                    surv = false;   // beyond the level-1 integer cut
                }
              }
              const unsigned m = __ballot_sync(0xffffffffu, surv);
              if (surv) wq[qn + __popc(m & ((1u << lane) - 1u))] = word;
              qn += __popc(m);
              __syncwarp();
            }
            if (qn == 0) { done = true; break; }   // nothing queued and no batch left
            const int take = qn < 32 ? qn : 32;
            if (lane < take) {
              const uint32_t word = wq[qn - take + lane];
              const uint32_t key = word >> cbits, k11 = word & cmask;
              const uint32_t b = n_pass > 1 ? key * n_pass + pass : key;
              const ColTerm ct = a.col_terms[b];
              const long long cb = ct.cb;
              const bool pos_side = (unsigned long long)ra * (unsigned long long)cb < (unsigned long long)k11 * (unsigned long long)N;
              const uint32_t kf = k11 < (uint32_t)kDomLevels ? k11 : (uint32_t)kDomLevels;
              ++evaluated_local;
              const long long k21 = cb - k11, k22 = N - ra - cb + k11;
              const double x11 = k11 < 32 ? x11tab[k11] : xlogx_u32(k11);
              const double x12 = k11 < kX12N ? x12tab[k11] : xlogx_u32((uint32_t)(ra - k11));
              const double x21 = k11 == 1 ? ct.x_cbm1 : xlogx_u32((uint32_t)k21), x22 = xlogx_u32((uint32_t)k22);
              double mat_e;
              if (varargs)
                mat_e = __dsub_rn(xN, __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(0.0, x11), x12), x21), x22));
              else
                mat_e = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(xN, x11), x12), x21), x22);
              const double sre = __dadd_rn(row_e, ct.col_e);
              const double v = (sre < mat_e) ? 0.0 : __dmul_rn(2.0, __dsub_rn(sre, mat_e));
              const bool min_ok = !a.has_min_llr || v >= a.min_llr;
              bool ok = v > 0.0 && min_ok;
              // (cells whose LLR rounds to 0 are cancellation noise: they teach nothing)
              bool strict_fail = v > 0.0 && !min_ok;
              const unsigned long long kbits = (unsigned long long)__double_as_longlong(v);
              e = make_uint4((uint32_t)kbits, (uint32_t)(kbits >> 32), b, k11);
              if (ok && vctrl[1]) {
                const uint4 thr = make_uint4((uint32_t)vctrl[4], (uint32_t)vctrl[5], (uint32_t)vctrl[6], (uint32_t)vctrl[7]);
                ok = !cand_better(thr, e);
                strict_fail = e.y < thr.y || (e.y == thr.y && e.x < thr.x);
              }
              if (strict_fail && pos_side)
                for (uint32_t kk = kf; kk >= 1 && (int)cb < vctrl[40 + kk]; --kk) atomicMin(&ctrl[40 + kk], (int)cb);
              pend_ok = ok;
            }
            qn -= take;
            __syncwarp();
          }
          // one meeting point per round.  Between this barrier and the next nobody appends, so n is group-uniform.
          const bool all_done = group_and<GROUP>(done);
          const int n_raw = vctrl[0];
          if (n_raw >= a.cbuf) {
            // the buffer filled up: keep between top_k and keep_max best, raise the threshold, make room
            if (GROUP > 32 && a.cbuf <= 512) {
              if (gw == 0) reduce_candidates<32>(tk, aux, a.cbuf, a.top_k, a.keep_max, hist, ctrl, lane);
            } else {
              reduce_candidates<GROUP>(tk, aux, a.cbuf, a.top_k, a.keep_max, hist, ctrl, gtid);
            }
          }
          if (all_done) break;
          group_sync<GROUP>();
        }
        group_sync<GROUP>();
      }
      if (!row_done) n_pass *= 2;
    }
    if (lane == 0) distinct_local += distinct_row;
    first_n = make_uint2(0u, 0u);
    if (NW == 1 && pf_first && item_n >= 0 && ub_n + lane < ue_n) first_n = a.ext[ub_n + lane];   // in flight during the final sort
    // ---- final select + write -------------------------------------------------------------------------------
    if (a.emit_all) {
      if (gtid == 0) a.out_len[item] = emitted;
    } else {
      int n = min(vctrl[0], a.cbuf);
      if (n > 0) {
        if (GROUP > 32 && n <= 512) {
          if (gw == 0) {
            int m = n;
            if (m > a.final_max) m = reduce_candidates<32>(tk, aux, m, a.top_k, a.final_max, hist, ctrl, lane);
            if (m <= 64 && !(a.tune & 8u)) sort_candidates_warp64(tk, m, lane); else sort_candidates<32>(tk, m, lane);
            if (lane == 0) ctrl[0] = m;
          }
          group_sync<GROUP>();
          n = vctrl[0];
        } else {
          if (n > a.final_max) n = reduce_candidates<GROUP>(tk, aux, n, a.top_k, a.final_max, hist, ctrl, gtid);
          if (GROUP == 32 && n <= 64 && !(a.tune & 8u)) sort_candidates_warp64(tk, n, lane); else sort_candidates<GROUP>(tk, n, gtid);
        }
        const int keep = n < a.top_k ? n : a.top_k;
        for (int i = gtid; i < keep; i += GROUP) {
          const size_t o = (size_t)item * a.out_stride + i;
          const uint4 c = tk[i];
          a.out_col[o] = (int32_t)c.z;
          a.out_llr[o] = __longlong_as_double((long long)(((unsigned long long)c.y << 32) | c.x));
          a.out_cnt[o] = (int32_t)c.w;
        }
        if (gtid == 0) a.out_len[item] = keep;
      } else if (gtid == 0) {
        a.out_len[item] = 0;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) distinct_local += __shfl_xor_sync(0xffffffffu, distinct_local, o);
  if (lane == 0 && distinct_local) atomicAdd(a.stat_distinct, distinct_local);
  for (int o = 16; o > 0; o >>= 1) evaluated_local += __shfl_xor_sync(0xffffffffu, evaluated_local, o);
  if (lane == 0 && evaluated_local) atomicAdd(a.stat_evaluated, evaluated_local);
}

}  // namespace cco
