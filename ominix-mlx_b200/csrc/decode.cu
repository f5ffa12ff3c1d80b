// decode.cu -- split-K GQA decode attention (Lq == 1), optionally fused with RoPE + KV append.
//
// Collapses the decode step of the reference's Attention::forward
// (qwen3-mlx/src/model.rs:186-212: rope(q, off), rope(k, off), cache.update_and_fetch, sdpa)
// into ONE launch.  HBM-bandwidth bound at batch sizes that fill the chip (every K/V byte is read exactly
// once); a chain of dependent round trips for a single sequence, which is what most of the structure below
// is about (profiles/r01_small_decode.md).
//
//   grid  = (num_splits, Hkv, B); one CTA streams a contiguous chunk of the keys of ONE
//           (batch, kv-head) and serves all G = Hq/Hkv query heads of that group from it.
//   16-bit, D = 128 ("hmma_tma"): a producer warp feeds a 3- or 6-stage ring of 64-key K/V tiles with
//           TMA (cp.async.bulk.tensor, 128 B swizzle, mbarrier completion, L2 evict-first); each consumer
//           warp owns whole tiles.  mma.sync m16n8k16 with SWAPPED roles -- keys / V features on the 16-row
//           M side, the query heads on the 8-wide N side, P^T through movmatrix -- so no row is padding.
//   float32 / other head dims ("simt"): 8 warps, 128-bit coalesced loads, FFMA dot products,
//           warp-shuffle reductions.
//   Prologue: all loads (q heads, rope row, norm weights, the new k / v rows) in ONE round trip, then
//           row statistics / RMSNorm / rotation from shared memory (stage_q, new_token).
//   The per-warp (m, l, O) states are merged in shared memory; with num_splits > 1 either the splits of a
//   (batch, kv-head) form a thread-block CLUSTER and are combined through distributed shared memory, or
//   partials go to a workspace and the LAST CTA of the pair (atomic ticket) combines them in one round trip
//   to L2 -- no second launch either way.  Counters reset themselves.
//   Fused mode: every CTA norms + ropes its G query heads on the fly (table lookup, reference
//   rounding); the CTA owning the last chunk also ropes k_new, stores k'/v_new into the cache
//   row (bit-identical to the unfused path) and folds that key in from shared memory, so the new row is
//   never re-read from HBM.
//   Variants on the same kernels: position read from device memory (CUDA-graph decode loop), head-sharded
//   output over NVLink peer stores, sequence-sharded float32 partials (+ seqshard_merge_kernel).
//   OMX_DECODE_TRACE=1|2 dumps per-CTA phase timelines (debugging aid; p.trace is null otherwise).
#include <algorithm>
#include <cstdio>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>
#include <cstdlib>
#include <vector>

#include "omx_common.cuh"
#include "omx_internal.h"
#include "sm100_utils.cuh"

// This file is compiled three times (csrc/Makefile) so that its ~50 kernels build in parallel:
//   OMX_DECODE_PART 0 (default): host logic and the TMA / tensor-core kernels
//   OMX_DECODE_PART 1 / 2 / 3: the bfloat16 / float16 / float32 instantiations of the CUDA-core kernel only
#ifndef OMX_DECODE_PART
#define OMX_DECODE_PART 0
#endif

// A/B switches for `make variant` (scripts/gpu_r02_variants.sh): which cold paths are kept out of line.
// One box, graph replay, us per fused step (all out of line / all inline / measured best = below):
//   Qwen3-0.6B bf16 ctx 2048 10.37 / 8.81 / 8.54, Qwen3-8B B1 ctx 8192 15.65 / 14.81 / 13.87,
//   C5 32 q / 8 kv ctx 32768 33.2 / 34.2 / 32.4, one rank of the sharded C5 14.60 / 15.15 / 14.55.
// The split-K combines stay inline (their arguments would travel through local memory); the array-mask score,
// the RMSNorm row sum and the new-token warp are calls.
#ifndef OMX_NI_COMBINE
#define OMX_NI_COMBINE __forceinline__
#endif
#ifndef OMX_NI_MASK
#define OMX_NI_MASK __noinline__
#endif
#ifndef OMX_NI_RMS
#define OMX_NI_RMS __noinline__
#endif
#ifndef OMX_NI_NT
#define OMX_NI_NT __noinline__
#endif

namespace omx {

namespace dd {  // shared by the three compilation parts of this file (same definition in each)

constexpr int kMaxPeers = OMX_MAX_PEERS;

// ---- data + flag ("LL") exchange over NVLink peer mappings
// staging of rank r: uint64 [2 (step parity)][world (source rank)][words]; word = {payload, sequence number}
struct LLDev {
  int world, rank;            // world == 0: off
  unsigned long long* buf[kMaxPeers];
  unsigned* seq;              // local count of completed steps
  int words;                  // words per source rank
};
__device__ __forceinline__ void st_ll(unsigned long long* a, uint32_t data, uint32_t flag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(a), "r"(data), "r"(flag) : "memory");
}
__device__ __forceinline__ uint4 ld_ll2(const void* a) {  // two adjacent words (16-byte aligned)
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_ll(const unsigned long long* a) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(a) : "memory");
  return v;
}
// One lane exchanges NW consecutive words starting at word w0 of its rank's slice: store to every peer, then
// poll the local staging buffer until every peer's words of step `seq` have landed; sink(src, words) consumes them.
template <int NW, typename Sink>
__device__ __forceinline__ void ll_exchange(const LLDev& ll, unsigned seq, int64_t w0, const uint32_t (&w)[NW],
                                            Sink&& sink) {
  const int64_t half = (int64_t)(seq & 1u) * ll.world;
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r) {
    if (r >= ll.world || r == ll.rank) continue;
    unsigned long long* dst = ll.buf[r] + (half + ll.rank) * ll.words + w0;
#pragma unroll
    for (int j = 0; j < NW; ++j) st_ll(dst + j, w[j], seq);
  }
  unsigned pend = ((1u << ll.world) - 1u) & ~(1u << ll.rank);
  const unsigned long long* mine = ll.buf[ll.rank] + half * ll.words + w0;
  unsigned spins = 0;
  while (pend) {
    uint2 v[kMaxPeers][NW];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if ((pend >> r) & 1u) {
#pragma unroll
        for (int j = 0; j < NW; ++j) v[r][j] = ld_ll(mine + (int64_t)r * ll.words + j);
      }
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if ((pend >> r) & 1u) {
        bool ok = true;
        uint32_t d[NW];
#pragma unroll
        for (int j = 0; j < NW; ++j) {
          ok = ok && v[r][j].y == seq;
          d[j] = v[r][j].x;
        }
        if (ok) {
          sink(r, d);
          pend &= ~(1u << r);
        }
      }
    if (++spins > (1u << 24)) __trap();  // a lost peer becomes a launch failure, not a hung GPU
  }
}
template <typename T>
struct LLPack;  // four consecutive elements <-> words
template <>
struct LLPack<float> {
  static constexpr int NW = 4;
  __device__ static void pack(const float4& v, uint32_t (&w)[4]) {
    w[0] = __float_as_uint(v.x); w[1] = __float_as_uint(v.y); w[2] = __float_as_uint(v.z); w[3] = __float_as_uint(v.w);
  }
  __device__ static void store(float* dst, const uint32_t (&w)[4]) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <typename T>
struct LLPack {  // 16-bit types
  static constexpr int NW = 2;
  __device__ static void pack(const float4& v, uint32_t (&w)[2]) {
    union { T t[4]; uint32_t u[2]; } x;
    x.t[0] = Num<T>::from_f(v.x); x.t[1] = Num<T>::from_f(v.y); x.t[2] = Num<T>::from_f(v.z); x.t[3] = Num<T>::from_f(v.w);
    w[0] = x.u[0]; w[1] = x.u[1];
  }
  __device__ static void store(T* dst, const uint32_t (&w)[2]) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
  }
};

struct DecodeParams {
  const void* q;
  void* out;
  int64_t qs[4], os[4];
  int B, Hq, Hkv, G;
  int Lk;     // keys attended (including the new row when fused)
  int n_mem;  // keys streamed from memory (Lk - 1 when fused)
  int D;
  float scale_log2;
  int num_splits, tiles_per_split;
  float* ws_o;   // [pair][split][G][D]
  float* ws_ml;  // [pair][split][G][2]
  int* counters; // [pair]
  // simt path reads K/V through plain pointers
  const void *k, *v;
  int64_t ks[4], vs[4];
  // fused new token
  int fused;
  const void *k_new, *v_new;
  int64_t kns[4], vns[4];
  void *k_row0, *v_row0;  // cache base pointers (row index Lk-1 is written)
  int64_t kcs[4], vcs[4];
  int rope_dims, traditional;
  const float *cos_row, *sin_row;  // table row of the current position, [rope_dims/2]
  // fused per-head RMSNorm of q / k_new before the rotation (null = none); [D] contiguous
  const void *q_norm_w, *k_norm_w;
  float norm_eps, norm_inv_n;
  // head-sharded output (C5): the final store goes to every rank's full buffer over NVLink peer
  // mappings; the last finishing CTA of the launch bumps one arrival counter per rank.
  int n_peers;                // 0: plain local store through `out`
  int peer_rank;
  int peer_total;             // CTAs that perform a final store in this launch
  void* peer_out[kMaxPeers];  // pre-offset to this rank's first global q head
  unsigned* peer_flag[kMaxPeers];  // rank r's counters live in rank r's memory: [world]
  int* peer_done;             // local, self-resetting
  int peer_wait;              // 1: the signalling CTA also waits for every peer's arrival (no separate wait launch)
  // array mask over the keys (unfused calls only): bool (true = keep) or additive in the q dtype,
  // broadcast [B,Hq,1,Lk] with element strides (batch, head, key)
  const void* mask;
  int mask_kind;  // 0 none, 1 bool, 2 additive
  int64_t mks[3];
  uint8_t* dead;  // [B][Hq]: 1 = the mask hid every key (masked_rows_fixup rewrites the row)
  // graph mode (omx_attn_decode_fused_dynamic): the position (= keys already cached) is read from device
  // memory at run time, so ONE captured launch serves every step of a decode loop.  Lk / n_mem /
  // tiles_per_split above then hold the values of the LAST admissible position (max_rows - 1) and
  // cos_row / sin_row the table base; the grid (num_splits) is fixed at capture.
  const int* pos_dev;
  int pos_stride;  // 0: one position shared by the batch; 1: pos_dev[b] (paged cache: per-sequence lengths)
  int max_rows;
  // paged KV (omx_attn_decode_fused_paged): rows live in a page pool [n_pages][Hkv][64][D] (one page = one
  // 64-key pipeline stage = one TMA box pair per tensor); tile t of sequence b is page
  // block_table[b * bt_stride + t]; k / v / k_row0 / v_row0 are the pool bases, ks[0] / vs[0] the PAGE strides.
  // Sequence b holds pos_dev[b] rows before the step (< 0: inactive slot, the CTA exits); the CTA that appends
  // kv head 0 stores pos + 1 to lens_out[b] (a second buffer: other CTAs of the launch still read pos_dev[b]).
  int paged;
  const int* block_table;
  int bt_stride;
  int* lens_out;
  // 1: the splits of one (batch, kv-head) form a thread-block cluster and are combined through
  // distributed shared memory (no partials in HBM/L2, no fence, no ticket)
  int cluster;
  // cluster combine, few splits (<= kPushSplits) and G <= 8: instead of publishing partials and PULLING column
  // slices from every peer (two cluster barriers), ranks 1.. PUSH their (O, m, l) into receive slots in rank 0's
  // shared memory and leave; rank 0 folds them after ONE cluster barrier.
  int push_combine;
  // sequence-sharded decode (omx_attn_decode_seqshard): `out` / peer_out are FLOAT32 partial slots
  // [B][Hq][D + 2] -- the normalised output of this rank's keys, then its (m, l) in the log2 domain -- and
  // os[] holds that slot's strides; append = 0: this rank attends but does not own the new token
  int partial;
  int append;
  // Rows [0, stable_rows) of the K/V buffers were written by launches that precede the stream's previous kernel
  // (the cache's bookkeeping: kv_cache_stable_rows): the TMA kernel may request tiles that lie wholly inside them
  // BEFORE the dependency wait of its programmatic launch.  0: nothing is read before the wait.
  int stable_rows;
  // TMA kernel: every tile request to shared memory also asks L2 for the tile `l2_ahead` tiles further on
  // (bytes in flight per SM beyond what the ring holds); l2_early = tiles past the ring requested into L2 before
  // the dependency wait.  0 = off.
  int l2_ahead, l2_early;
  // 1: split-K combine by ALL CTAs of the pair (one-wave grids only: every CTA is resident).  Each CTA publishes
  // its partial, the CTAs of a (batch, kv-head) pair meet at a counter, and every CTA then folds ITS slice of the
  // output columns over all partials in one round trip to L2 -- instead of the last CTA folding everything
  // (a 64-way split: 6.7 us of dependent round trips in one CTA).  counters2[pair] counts the CTAs that have left.
  int gsync;          // 1: meet at a counter; 2: no meeting point -- partials travel as {value, launch tag} words
  int* counters2;
  // gsync == 2: every float of a partial (and its m, l) is an 8-byte word {value, tag}, tag = *gs_seq + 1 = this
  // launch's number; readers poll the words they need, so the writer needs no fence and nobody takes a counter
  // round trip.  ws_w: [pair][split][head][D + 2] words in a pool that only ever holds such words (zeroed when
  // allocated; tags only grow), gs_seq: launches completed on that pool, bumped by the launch's last CTA.
  uint2* ws_w;
  unsigned* gs_seq;
  // CUDA-core kernel, staged variant: byte offset (from the dynamic shared memory base) of the K stage, and the
  // keys one stage holds (V follows K); 0 = rows are loaded straight into registers
  int stage_off, stage_keys;
  // data + flag exchange of the head-sharded step (omx_attn_decode_fused_sharded_ll; all-CTA combine only): the
  // lane that holds four final output values stores them as {payload, sequence number} words into every peer's
  // staging buffer, then polls its own staging buffer for the peers' words and unpacks them into ll_out (the
  // local full-head output, strides os[]).  n_peers stays 0: `out` is the private local slice.
  LLDev ll;
  void* ll_out;
  // debugging aid (OMX_DECODE_TRACE=1): per-CTA phase timestamps, [cta][16] x %globaltimer ns; null otherwise
  unsigned long long* trace;
};

// 16-bit instantiations of the CUDA-core kernel (parts 1 / 2), called from decode_attention (part 0)
void launch_simt_bf16(DecodeParams& p, cudaStream_t stream, int Gt, dim3 grid, bool one_wave, bool want_cluster);
void launch_simt_f16(DecodeParams& p, cudaStream_t stream, int Gt, dim3 grid, bool one_wave, bool want_cluster);
void launch_simt_f32(DecodeParams& p, cudaStream_t stream, int Gt, dim3 grid, bool one_wave, bool want_cluster);

}  // namespace dd

namespace {

using dd::DecodeParams;
using dd::kMaxPeers;

constexpr int kTile = 64;              // keys per pipeline stage
constexpr int kBoxBytes = 64 * 64 * 2;  // one TMA box: 64 keys x 64 features x 2 B
constexpr int kStageBytes = 4 * kBoxBytes;  // K lo/hi + V lo/hi
constexpr int kQPitch = 136;           // padded q row (elements) -> conflict-free fragment loads
constexpr int kMaxSplits = 64;         // split-K upper bound (plan_splits), sizes the combine scratch


__device__ __forceinline__ void trace_mark(const DecodeParams& p, int slot) {
  if (p.trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    p.trace[(size_t)cta * 16 + slot] = t;
  }
}

// Programmatic dependent launch (PDL).  Decode launches carry the programmatic-stream-serialization attribute:
// their CTAs may be dispatched while the previous kernel on the stream is still running, and wait -- before
// their first access to memory that kernel may write (q / k_new / v_new produced by the caller's projection,
// the cache row it appended, split-K scratch, counters) -- until it has completed and its writes are visible.
// The TMA kernel releases the NEXT grid right after its own wait: a grid is released once every CTA of its
// predecessor has said so or exited, so at most two consecutive launches overlap and everything written two
// launches back is complete.  What the overlap buys: the next launch's CTAs take the SMs this launch leaves
// idle (a 128-CTA grid on 148 SMs, CTAs that finish early) and fill their TMA rings with cache rows that are
// older than the previous launch (DecodeParams::stable_rows) while this launch drains its slowest CTAs and
// combines; launch latency, barrier init and the tensor-map fetch go under the same cover.  A predecessor that
// never triggers releases its dependents on completion, i.e. behaves like a plain stream-ordered launch.
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release_next_grid() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// The launch's key count, split size and rope row: kernel arguments, or derived from *pos_dev.
struct DecodeDyn {
  int Lk, n_mem, tps;
  const float *cos_row, *sin_row;
};

__device__ __forceinline__ DecodeDyn load_dyn(const DecodeParams& p, int b) {
  DecodeDyn d{p.Lk, p.n_mem, p.tiles_per_split, p.cos_row, p.sin_row};
  if (p.pos_dev) {
    // clamped: a loop driven past the reserved rows rewrites the last row instead of leaving the buffer
    // (omx_kv_cache_advance reports the overflow to the host)
    const int pos = min(max(__ldg(p.pos_dev + (size_t)b * p.pos_stride), 0), p.max_rows - 1);
    d.n_mem = pos;
    d.Lk = pos + 1;
    const int n_tiles = (pos + kTile - 1) / kTile;
    d.tps = max(1, (n_tiles + p.num_splits - 1) / p.num_splits);
    if (p.rope_dims > 0) {
      d.cos_row = p.cos_row + (size_t)pos * (p.rope_dims >> 1);
      d.sin_row = p.sin_row + (size_t)pos * (p.rope_dims >> 1);
    }
  }
  return d;
}

// score (log2 domain) of `key` for query head `h` after the array mask.  Additive entries <= -1e8
// (the callers' -1e9 / -inf spelling of "hidden") hide the key outright, like the prefill kernel.
template <typename T>
__device__ OMX_NI_MASK float mask_score(const DecodeParams& p, float s, int b, int h, int key) {
  const int64_t mi = b * p.mks[0] + h * p.mks[1] + key * p.mks[2];
  if (p.mask_kind == 1) return ((const uint8_t*)p.mask)[mi] ? s : -INFINITY;
  const float mf = Num<T>::to_f(((const T*)p.mask)[mi]);
  return mf <= -1e8f ? -INFINITY : fmaf(mf, kLog2e, s);
}

// Code size matters here: a single-sequence launch runs every instruction of its tail once per CTA, on a cold
// instruction cache (the kernels were 560 KB of SASS with the peer loops unrolled into each of the ~20 final
// store sites).  The common case -- a local store in the output type -- stays inline; everything else is a call.
template <typename T>
__device__ __noinline__ void store_out_slow(const DecodeParams& p, int64_t off, float v) {
  if (p.partial) {  // f32 partial slot(s)
    if (p.n_peers == 0) ((float*)p.out)[off] = v;
#pragma unroll 1
    for (int r = 0; r < p.n_peers; ++r) ((float*)p.peer_out[r])[off] = v;
    return;
  }
  const T x = Num<T>::from_f(v);
#pragma unroll 1
  for (int r = 0; r < p.n_peers; ++r) ((T*)p.peer_out[r])[off] = x;
}
template <typename T>
__device__ __forceinline__ void store_out(const DecodeParams& p, int64_t off, float v) {
  if (!p.partial && p.n_peers == 0) {
    ((T*)p.out)[off] = Num<T>::from_f(v);
    return;
  }
  store_out_slow<T>(p, off, v);
}

// (m, l) of one (batch, head) behind its D partial-output floats (sequence-sharded mode only)
__device__ __noinline__ void store_ml_slow(const DecodeParams& p, int b, int head, float M, float L) {
  const int64_t off = b * p.os[0] + (int64_t)head * p.os[1] + (int64_t)p.D * p.os[3];
  if (p.n_peers == 0) {
    ((float*)p.out)[off] = M;
    ((float*)p.out)[off + 1] = L;
  }
#pragma unroll 1
  for (int r = 0; r < p.n_peers; ++r) {
    ((float*)p.peer_out[r])[off] = M;
    ((float*)p.peer_out[r])[off + 1] = L;
  }
}
__device__ __forceinline__ void store_ml(const DecodeParams& p, int b, int head, float M, float L) {
  if (p.partial) store_ml_slow(p, b, head, M, L);
}

// After the final stores of one CTA (called by all its threads).  Writers fence their peer stores
// at system scope, the CTA takes a ticket, and the launch's last ticket publishes the arrival.
__device__ __noinline__ void peer_signal_slow(const DecodeParams& p, int tid);
__device__ __forceinline__ void peer_signal(const DecodeParams& p, int tid) {
  if (p.n_peers) peer_signal_slow(p, tid);
}
__device__ __noinline__ void peer_signal_slow(const DecodeParams& p, int tid) {
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    const int t = atomicAdd(p.peer_done, 1);
    if (t == p.peer_total - 1) {
      *p.peer_done = 0;  // self-reset for the next launch (stream-ordered)
      __threadfence_system();
      unsigned mine = 0;
      for (int r = 0; r < p.n_peers; ++r) {
        if (r == p.peer_rank) mine = atomicAdd_system(p.peer_flag[r] + p.peer_rank, 1u) + 1u;
        else atomicAdd_system(p.peer_flag[r] + p.peer_rank, 1u);
      }
      if (p.peer_wait) {
        // The launch ends only when every rank's slice of THIS step has landed in the local buffer: the arrival
        // counters only grow, "this step" = as many arrivals as this rank has itself signalled.  Replaces the
        // separate one-warp wait launch (a kernel boundary + its launch latency per step).
        const volatile unsigned* f = p.peer_flag[p.peer_rank];
        for (int r = 0; r < p.n_peers; ++r) {
          unsigned spins = 0;
          while ((int)(f[r] - mine) < 0) {
            __nanosleep(32);
            if (++spins > (1u << 25)) __trap();  // a lost peer becomes a launch failure, not a hung GPU
          }
        }
        __threadfence_system();
      }
    }
  }
}

// ---- prologue: everything the CTA needs besides K/V arrives in ONE parallel round trip ----
// Single-sequence decode is a chain of dependent memory round trips (per-CTA timelines, OMX_DECODE_TRACE):
// the earlier prologue spent 2-5 us in front of the first MMA on q -> row statistics -> weights / rope row,
// each step waiting on global memory.  Now every thread first issues its share of ALL the loads (raw q
// heads, rope row, norm weights and -- in the CTA that appends -- the new k / v rows), the CTA syncs once,
// and the row statistics, normalisation and rotation run from shared memory only.
// named barriers (id 1..15; id 0 is __syncthreads): `count` threads must arrive, sync blocks, arrive does not
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// 1 / rms of a row held in shared memory: the reference's strict left-to-right f32 sum of squares.  The
// row is pulled into registers with 128-bit loads first, so what remains serial is the FADD chain alone
// (the scalar loop took 2.7 us for 128 bf16 elements: one exposed LDS per step).
template <typename E>
__device__ OMX_NI_RMS float rms_rsqrt_smem(const E* row, int D, float eps, float inv_n) {
  constexpr int V = 16 / (int)sizeof(E);
  float acc = 0.f;
  if ((D % (4 * V)) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    for (int c = 0; c < D; c += 4 * V) {
      union { uint4 r[4]; E t[4 * V]; } u;
#pragma unroll
      for (int i = 0; i < 4; ++i) u.r[i] = *reinterpret_cast<const uint4*>(row + c + i * V);
      // squares first (independent, pipelined), then the serial FADD chain: same operations and order as
      // acc = fadd(acc, fmul(v, v)) per element, but the chain is 4 cycles per element instead of
      // convert -> FMUL -> FADD (traced: 1.2-2 us for a 128-element row before)
      float sq[4 * V];
#pragma unroll
      for (int i = 0; i < 4 * V; ++i) {
        const float v = Num<E>::to_f(u.t[i]);
        sq[i] = __fmul_rn(v, v);
      }
#pragma unroll
      for (int i = 0; i < 4 * V; ++i) acc = __fadd_rn(acc, sq[i]);
    }
  } else {
    for (int d = 0; d < D; ++d) {
      const float v = Num<E>::to_f(row[d]);
      acc = __fadd_rn(acc, __fmul_rn(v, v));
    }
  }
  return __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fmul_rn(acc, inv_n), eps)));
}

struct PrologueSmem {       // static shared memory of both kernels
  float cs[256];            // cos row | sin row, rounded to T (rope_dims <= 256)
  float qw[256], kw[256];   // q_norm / k_norm weights as float (head_dim <= 256)
  float rs[16];             // 1/rms of the staged q heads
};

// after_loads(): called once the prologue's own loads are issued and before anything waits on them (the
// CUDA-core kernel requests its first K/V rows there).
// sync(): barrier over the `nthr` threads that run the prologue (the whole CTA, or the consumer warps only).
template <typename T, typename F, typename S>
__device__ __forceinline__ void stage_q(const DecodeParams& p, T* q_s, int pitch, int rows_total, int first_head,
                                        int n_heads, int b, int hk, int tid, int nthr, PrologueSmem& ps,
                                        float* nt_k, float* nt_v, bool has_nt, const DecodeDyn& dy,
                                        F&& after_loads, S&& sync) {
  const int D = p.D;
  const bool rope = p.fused && p.rope_dims > 0;
  const int half = p.rope_dims >> 1;
  const T* qg = (const T*)p.q + b * p.qs[0] + (int64_t)first_head * p.qs[1];
  // ---- loads (all independent)
  constexpr int V = 16 / (int)sizeof(T);
  const bool vec = p.qs[3] == 1 && (D % V) == 0 && (pitch % V) == 0 && p.qs[1] % V == 0 &&
                   ((reinterpret_cast<uintptr_t>(qg) | reinterpret_cast<uintptr_t>(q_s)) & 15) == 0;
  if (vec) {
    const int per = D / V;
    for (int idx = tid; idx < n_heads * per; idx += nthr) {
      const int g = idx / per, c = idx % per;
      *reinterpret_cast<uint4*>(q_s + g * pitch + c * V) = *reinterpret_cast<const uint4*>(qg + g * p.qs[1] + c * V);
    }
  } else {
    for (int idx = tid; idx < n_heads * D; idx += nthr) {
      const int g = idx / D, d = idx % D;
      q_s[g * pitch + d] = qg[g * p.qs[1] + d * p.qs[3]];
    }
  }
  for (int idx = tid; idx < (rows_total - n_heads) * D; idx += nthr)
    q_s[(n_heads + idx / D) * pitch + idx % D] = Num<T>::from_f(0.f);
  const int rt = nthr - 1 - tid;  // the small tables start from the other end of the CTA
  if (rope)
    for (int u = rt; u < half; u += nthr) {
      ps.cs[u] = rnd<T>(dy.cos_row[u]);
      ps.cs[half + u] = rnd<T>(dy.sin_row[u]);
    }
  if (p.q_norm_w)
    for (int d = rt; d < D; d += nthr) ps.qw[d] = Num<T>::to_f(((const T*)p.q_norm_w)[d]);
  if (has_nt) {
    const T* kn = (const T*)p.k_new + b * p.kns[0] + hk * p.kns[1];
    const T* vn = (const T*)p.v_new + b * p.vns[0] + hk * p.vns[1];
    for (int d = rt; d < D; d += nthr) {
      nt_k[d] = Num<T>::to_f(kn[d * p.kns[3]]);
      nt_v[d] = Num<T>::to_f(vn[d * p.vns[3]]);
    }
    if (p.k_norm_w)
      for (int d = rt; d < D; d += nthr) ps.kw[d] = Num<T>::to_f(((const T*)p.k_norm_w)[d]);
  }
  after_loads();
  trace_mark(p, 14);
  if (!rope && !p.q_norm_w) return;  // the caller's barrier publishes the copy
  sync();
  trace_mark(p, 12);
  // ---- q_norm: the reference's left-to-right f32 sum, one thread per head, from shared memory
  if (p.q_norm_w) {
    // spread over the warps (one head per warp's lane 0) so the serial chains run on different schedulers
    if ((tid & 31) == 0 && (tid >> 5) < n_heads)
      ps.rs[tid >> 5] = rms_rsqrt_smem<T>(q_s + (tid >> 5) * pitch, D, p.norm_eps, p.norm_inv_n);
    for (int g = (nthr >> 5) + tid; g < n_heads; g += nthr)  // more heads than warps (G = 16 on 4 warps)
      ps.rs[g] = rms_rsqrt_smem<T>(q_s + g * pitch, D, p.norm_eps, p.norm_inv_n);
    sync();
    trace_mark(p, 13);
  }
  // ---- normalise + rotate in place (every element is owned by exactly one thread)
  const bool nrm = p.q_norm_w != nullptr;
  auto qval = [&](int g, int d) -> float {
    const float x = Num<T>::to_f(q_s[g * pitch + d]);
    return nrm ? rms_apply<T>(x, ps.rs[g], ps.qw[d], true) : x;
  };
  if (rope) {
    const int per = half + (D - p.rope_dims);
    for (int idx = tid; idx < n_heads * per; idx += nthr) {
      const int g = idx / per, u = idx % per;
      if (u < half) {
        const int i1 = p.traditional ? 2 * u : u;
        const int i2 = p.traditional ? 2 * u + 1 : u + half;
        float o1, o2;
        rope_pair<T>(qval(g, i1), qval(g, i2), ps.cs[u], ps.cs[half + u], o1, o2);
        q_s[g * pitch + i1] = Num<T>::from_f(o1);
        q_s[g * pitch + i2] = Num<T>::from_f(o2);
      } else if (nrm) {
        const int d = p.rope_dims + (u - half);
        q_s[g * pitch + d] = Num<T>::from_f(qval(g, d));
      }
    }
  } else {
    for (int idx = tid; idx < n_heads * D; idx += nthr) {
      const int g = idx / D, d = idx % D;
      q_s[g * pitch + d] = Num<T>::from_f(qval(g, d));
    }
  }
}

// One warp: norm + rope the new k row (staged raw in nt_k by stage_q, v in nt_v), append k'/v_new to the
// cache row, and score the new key against the staged q heads -- from shared memory only.
// nt_k/nt_v: float[D], nt_m[g] <- log2-domain score.
template <typename T>
__device__ OMX_NI_NT void new_token(const DecodeParams& p, const T* q_s, int pitch, int n_heads,
                                          int b, int hk, int lane, float* nt_k, float* nt_v,
                                          float* nt_m, const PrologueSmem& ps, const DecodeDyn& dy,
                                          bool write_cache = true) {
  const int D = p.D;
  const bool nrm = p.k_norm_w != nullptr;
  float kr = 0.f;
  if (nrm) {  // left-to-right f32 sum over the raw row
    if (lane == 0) kr = rms_rsqrt_smem<float>(nt_k, D, p.norm_eps, p.norm_inv_n);
    kr = __shfl_sync(0xffffffffu, kr, 0);
  }
  // k_new element after the optional RMSNorm (rounded to T); nt_k[d] holds the raw value until the lane
  // that owns it overwrites it below
  auto kval = [&](int d) -> float { return nrm ? rms_apply<T>(nt_k[d], kr, ps.kw[d], true) : nt_k[d]; };
  // cache row of the new token: row Lk - 1 of sequence b, or (paged) row (Lk - 1) % 64 of its last page
  int64_t cb = b, crow = dy.Lk - 1;
  if (p.paged) {
    cb = __ldg(p.block_table + (size_t)b * p.bt_stride + (crow >> 6));
    crow &= 63;
  }
  T* kc = (T*)p.k_row0 + cb * p.kcs[0] + hk * p.kcs[1] + crow * p.kcs[2];
  T* vc = (T*)p.v_row0 + cb * p.vcs[0] + hk * p.vcs[1] + crow * p.vcs[2];
  const int half = p.rope_dims >> 1;
  for (int u = lane; u < half; u += 32) {
    const int i1 = p.traditional ? 2 * u : u;
    const int i2 = p.traditional ? 2 * u + 1 : u + half;
    float o1, o2;
    rope_pair<T>(kval(i1), kval(i2), ps.cs[u], ps.cs[half + u], o1, o2);
    if (write_cache) {
      kc[i1 * p.kcs[3]] = Num<T>::from_f(o1);
      kc[i2 * p.kcs[3]] = Num<T>::from_f(o2);
    }
    nt_k[i1] = o1;
    nt_k[i2] = o2;
  }
  for (int d = p.rope_dims + lane; d < D; d += 32) {
    const float x = kval(d);
    if (write_cache) kc[d * p.kcs[3]] = Num<T>::from_f(x);
    nt_k[d] = x;
  }
  if (write_cache)
    for (int d = lane; d < D; d += 32) vc[d * p.vcs[3]] = Num<T>::from_f(nt_v[d]);  // exact round trip
  __syncwarp();
  for (int g = 0; g < n_heads; ++g) {
    float a = 0.f;
    for (int d = lane; d < D; d += 32) a = fmaf(Num<T>::to_f(q_s[g * pitch + d]), nt_k[d], a);
    a = warp_sum(a);
    if (lane == 0) nt_m[g] = a * p.scale_log2;
  }
}

// all-CTA combine: NS splits are laid over groups of W lanes, K groups per output column
struct GsyncShape {
  int W, K;
};
__host__ __device__ inline GsyncShape gsync_shape(int ns) {
  if (ns > 32) return {32, (ns + 31) / 32};
  int w = 1;
  while (w < ns) w <<= 1;
  return {w, 1};
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- thread-block cluster primitives (split-K combine through distributed shared memory)
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// "every CTA of the cluster has started": arrive at kernel entry, wait before the first store into a peer's shared
// memory (the push combine stores before any other cluster barrier; compute-sanitizer flags a DSMEM write to a block
// that "might not have entered yet" otherwise)
__device__ __forceinline__ void cluster_arrive_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_plain() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t dsmem_addr(const void* local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_dsmem_f32(uint32_t a, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
constexpr int kPushSplits = 4;  // receive slots are sized (splits - 1) x G x (D + 2) floats per CTA
__host__ __device__ constexpr int push_recv_floats(int splits, int heads, int D) { return (splits - 1) * heads * (D + 2); }
__device__ __forceinline__ float2 ld_dsmem_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
constexpr int kMaxClusterSplits = 16;  // non-portable cluster size limit on sm_100
// floats of cluster scratch behind the merge inputs: partial (m, l) | weights | sums, per row pitch
__host__ __device__ constexpr int cluster_scratch_floats(int rows) { return (2 + 2 * kMaxClusterSplits) * rows; }

// gsync == 2: a CTA that leaves without taking part (inactive paged slot) still counts towards "the launch is over",
// or the launch's tag would never be retired
__device__ __forceinline__ void gs_leave_unused(const DecodeParams& p, unsigned gs_tag, int tid) {
  if (p.gsync == 2 && tid == 0 && atomicAdd(p.peer_done, 1) == p.peer_total - 1) {
    *p.peer_done = 0;
    *p.gs_seq = gs_tag;
  }
}

// ---- all-CTA split-K combine (DecodeParams::gsync; one-wave grids).  publish -> meet -> every CTA folds its own
// slice of the output columns.  Lanes run over the SPLITS: a group of W lanes holds one float4 column's
// contributions, so the per-head maximum / sum and the fold are warp shuffles; what is left of the tail is one
// round trip to L2 behind the meeting point (partials and (m, l) requested together) and at most one CTA barrier.
// Kept out of line and un-unrolled on purpose (see store_out_slow).
template <typename T>
__device__ __noinline__ void ll_quad(const DecodeParams& p, unsigned seq, int b, int head, int d, float4 v) {
  using LP = dd::LLPack<T>;
  uint32_t w[LP::NW];
  LP::pack(v, w);
  LP::store((T*)p.out + b * p.os[0] + (int64_t)head * p.os[1] + d, w);  // the local slice (host: os[3] == 1)
  const int64_t e = ((int64_t)b * p.Hq + head) * p.D + d;              // element of this rank's [B,Hq,D] slice
  dd::ll_exchange<LP::NW>(p.ll, seq, e * (int64_t)sizeof(T) / 4, w, [&](int src, const uint32_t (&r)[LP::NW]) {
    LP::store((T*)p.ll_out + b * p.os[0] + ((int64_t)src * p.Hq + head) * p.os[1] + d, r);
  });
}

template <typename T>
__device__ __noinline__ void gsync_combine(const DecodeParams& p, float* scratch, int first_head, int n_heads,
                                           int b, int pair, int split, int tid, int nthr, unsigned gs_tag) {
  const int NS = p.num_splits, D = p.D;
  const bool tagged = p.gsync == 2;
  // (A warm-up walk of this function by the idle producer warp while the tiles are in flight -- every load issued,
  // no store / barrier / atomic -- was tried against the instruction-fetch cost of the tail and measured SLOWER:
  // one rank of the sharded C5 14.0 -> 17.0 us, the walk outlasts a 4-tile key loop and the phases below did not
  // get shorter; scripts/gpu_r02_warm_ab.sh.)
  if (!tagged) __threadfence();
  __syncthreads();  // (tagged: the merge inputs in `scratch` are dead from here on)
  trace_mark(p, 4);
  const int D4 = D >> 2;
  const int d4sh = 31 - __clz(D4);            // head dims are powers of two (decode_supported): col / D4 = col >> d4sh
  const int C = n_heads * D4;                 // float4 columns of the pair's output
  const int slice = (C + NS - 1) / NS;
  const int c0 = split * slice, c1 = min(C, c0 + slice);
  const int ncol = max(0, c1 - c0);
  const GsyncShape gs = gsync_shape(NS);
  const int W = gs.W, K = gs.K;
  const int wsh = 31 - __clz(W);              // W = 1 << wsh
  const int n_wi = W == 32 ? ncol * K : (ncol + (32 >> wsh) - 1) >> (5 - wsh);  // warp-sized work items
  const int nwarps = nthr >> 5, warp = tid >> 5, lane = tid & 31;
  const int g0 = ncol ? c0 >> d4sh : 0, g1 = ncol ? (c1 - 1) >> d4sh : -1;  // heads the slice touches
  const int nh = g1 - g0 + 1;
  const int mlp = (NS * nh + 3) & ~3;
  float* sm_m = scratch;                      // [NS][nh]  (the merge inputs are dead: the own partial is in ws)
  float* sm_l = sm_m + mlp;                   // [NS][nh]
  float4* red = reinterpret_cast<float4*>(sm_l + mlp);      // [ncol][K]   (K > 1 only)
  float* sm_inv = reinterpret_cast<float*>(red + ncol * K);  // [ncol]
  const int64_t e0p = (int64_t)pair * NS * n_heads;
  const int64_t ob = b * p.os[0];
  if (!tagged) {
    if (tid == 0) {
      atomicAdd(&p.counters[pair], 1);
      unsigned spins = 0;
      while (ld_acquire_gpu(&p.counters[pair]) < NS)
        if (++spins > (1u << 26)) __trap();  // a CTA that never became resident: a launch failure, not a hung GPU
    }
    __syncthreads();
  }
  trace_mark(p, 5);
  // tagged words: poll until both words of a 16-byte pair carry this launch's tag (bounded: see above)
  auto poll2 = [&](const uint2* w) {
    unsigned spins = 0;
    uint4 v = dd::ld_ll2(w);
    while (v.y != gs_tag || v.w != gs_tag) {
      if (++spins > (1u << 24)) __trap();
      __nanosleep(128);  // thousands of lanes poll: leave the L2 to the CTAs that are still streaming their keys
      v = dd::ld_ll2(w);
    }
    return make_float2(__uint_as_float(v.x), __uint_as_float(v.z));
  };
  const unsigned ll_seq = p.ll.world ? __ldcg(p.ll.seq) + 1u : 0u;
  // work item wi of this warp: local column colL, split sp of this lane
  auto item = [&](int wi, int& colL, int& sp) {
    if (W == 32) {
      colL = wi / K;
      sp = (wi - colL * K) * 32 + lane;
    } else {
      colL = (wi << (5 - wsh)) + (lane >> wsh);
      sp = lane & (W - 1);
    }
    return wi < n_wi && colL < ncol && sp < NS;
  };
  auto fetch = [&](int wi) {
    int colL, sp;
    if (!item(wi, colL, sp)) return make_float4(0.f, 0.f, 0.f, 0.f);
    const int col = c0 + colL;
    if (tagged) {
      const uint2* w = p.ws_w + (e0p + (int64_t)sp * n_heads + (col >> d4sh)) * (D + 2) + (col & (D4 - 1)) * 4;
      const float2 lo = poll2(w), hi = poll2(w + 2);
      return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    return __ldcg(reinterpret_cast<const float4*>(p.ws_o + (e0p + (int64_t)sp * n_heads + (col >> d4sh)) * D + (col & (D4 - 1)) * 4));
  };
  float4 cur = fetch(warp);
#pragma unroll 1
  for (int idx = tid; idx < NS * nh; idx += nthr) {
    const int sp = nh == 1 ? idx : idx / nh, g = g0 + idx - sp * nh;
    const float2 ml = tagged ? poll2(p.ws_w + (e0p + (int64_t)sp * n_heads + g) * (D + 2) + D)
                             : __ldcg(reinterpret_cast<const float2*>(&p.ws_ml[(e0p + (int64_t)sp * n_heads + g) * 2]));
    sm_m[idx] = ml.x;
    sm_l[idx] = ml.y;
  }
  __syncthreads();
  trace_mark(p, 9);
  auto emit4 = [&](int col, const float4& v) {
    const int g = col >> d4sh, d = (col & (D4 - 1)) * 4;
    if (p.ll.world) {
      ll_quad<T>(p, ll_seq, b, first_head + g, d, v);
      return;
    }
    const int64_t o = ob + (int64_t)(first_head + g) * p.os[1] + (int64_t)d * p.os[3];
    store_out<T>(p, o, v.x);
    store_out<T>(p, o + p.os[3], v.y);
    store_out<T>(p, o + 2 * p.os[3], v.z);
    store_out<T>(p, o + 3 * p.os[3], v.w);
  };
#pragma unroll 1
  for (int wi = warp; wi < n_wi; wi += nwarps) {  // (warp-uniform)
    const float4 nxt = fetch(wi + nwarps);        // the next item's partial is in flight under this one's fold
    int colL, sp;
    const bool valid = item(wi, colL, sp);
    const int cc = min(colL, ncol - 1);           // lanes past the slice compute on its last column and store nothing
    const int col = c0 + cc, g = col >> d4sh, gi = g - g0;
    const int l0 = lane & (W - 1);
    float M = -INFINITY;
#pragma unroll 1
    for (int s2 = l0; s2 < NS; s2 += W) M = fmaxf(M, sm_m[s2 * nh + gi]);
#pragma unroll 1
    for (int o = W >> 1; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    float L = 0.f;
#pragma unroll 1
    for (int s2 = l0; s2 < NS; s2 += W) {
      const float ms = sm_m[s2 * nh + gi];
      if (ms > -INFINITY) L = fmaf(sm_l[s2 * nh + gi], fast_exp2(ms - M), L);
    }
    const float mine = valid ? sm_m[sp * nh + gi] : -INFINITY;
    const float e = mine > -INFINITY ? fast_exp2(mine - M) : 0.f;
    float4 c = make_float4(cur.x * e, cur.y * e, cur.z * e, cur.w * e);
#pragma unroll 1
    for (int o = W >> 1; o > 0; o >>= 1) {
      L += __shfl_xor_sync(0xffffffffu, L, o);
      c.x += __shfl_xor_sync(0xffffffffu, c.x, o);
      c.y += __shfl_xor_sync(0xffffffffu, c.y, o);
      c.z += __shfl_xor_sync(0xffffffffu, c.z, o);
      c.w += __shfl_xor_sync(0xffffffffu, c.w, o);
    }
    const int kk = W == 32 ? wi - (wi / K) * K : 0;  // which 32-split block of the column
    const bool lead = colL < ncol && l0 == 0 && kk == 0;
    if (lead && (col & (D4 - 1)) == 0) {  // the CTA that owns the head's first column reports the row
      store_ml(p, b, first_head + g, M, L);
      if (p.dead) p.dead[(int64_t)b * p.Hq + first_head + g] = L > 0.f ? 0 : 1;
    }
    if (K == 1) {
      if (lead) {
        const float inv = 1.0f / L;
        emit4(col, make_float4(c.x * inv, c.y * inv, c.z * inv, c.w * inv));
      }
    } else if (colL < ncol && lane == 0) {
      red[colL * K + kk] = c;
      if (lead) sm_inv[colL] = 1.0f / L;
    }
    cur = nxt;
  }
  if (K > 1) {
    __syncthreads();
    trace_mark(p, 10);
#pragma unroll 1
    for (int colL = tid; colL < ncol; colL += nthr) {
      float4 a = red[colL * K];
#pragma unroll 1
      for (int k = 1; k < K; ++k) {
        const float4 r = red[colL * K + k];
        a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
      }
      const float inv = sm_inv[colL];
      emit4(c0 + colL, make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv));
    }
  }
  trace_mark(p, 11);
  peer_signal(p, tid);  // (peer_total counts every CTA of the launch in this mode)
  if (p.ll.world) __syncthreads();  // this CTA's exchange is complete
  if (tid == 0) {
    if (!tagged) {
      const int t = atomicAdd(&p.counters2[pair], 1);
      if (t == NS - 1) {  // everybody has passed the meeting point: reset both counters for the next launch
        p.counters[pair] = 0;
        p.counters2[pair] = 0;
      }
    }
    if ((tagged || p.ll.world) && atomicAdd(p.peer_done, 1) == p.peer_total - 1) {
      *p.peer_done = 0;     // the launch's last CTA: every word of the step has been sent and received here
      if (p.ll.world) *p.ll.seq = ll_seq;
      if (tagged) *p.gs_seq = gs_tag;
    }
  }
  trace_mark(p, 6);
}

// The combines below run once per CTA at the very end of the launch; each is a function of its own so that a
// launch only walks the code of the one it uses (see store_out_slow on code size).
struct MergeArgs {
  const float* mo;
  int n_ent, rows, first_head, n_heads, b, pair, split, tid, nthr;
  int* s_ticket;
  float *cl, *recv;
};

template <typename T>
__device__ OMX_NI_COMBINE void combine_push(const DecodeParams& p, const MergeArgs& a) {
  const float* mo = a.mo;
  const int n_ent = a.n_ent, rows = a.rows, first_head = a.first_head, n_heads = a.n_heads, b = a.b, pair = a.pair,
            split = a.split, tid = a.tid, nthr = a.nthr;
  int* s_ticket = a.s_ticket;
  float *cl = a.cl, *recv = a.recv;
  const int D = p.D;
  const int64_t ob = b * p.os[0];
  float* part_o = const_cast<float*>(mo);
  auto recv_at = [&](int sp, int g, int d) { return recv + ((sp - 1) * n_heads + g) * (D + 2) + d; };
  (void)n_ent; (void)rows; (void)pair; (void)split; (void)s_ticket; (void)cl; (void)recv; (void)ob; (void)part_o; (void)recv_at;
    // ---- push combine: ranks 1.. have stored into rank 0's receive slots; their arrive (release) publishes the
    // stores and they leave -- nobody reads THEIR shared memory.  Rank 0 folds after the one barrier; every
    // element is owned by the thread that produced rank 0's own partial, so no further CTA barrier is needed.
    cluster_arrive_release();
    if (split != 0) {
      trace_mark(p, 6);
      return;
    }
    cluster_wait_acquire();
    __syncthreads();  // rank 0's own (m, l) in `cl` were written by the d == 0 owners only
    trace_mark(p, 4);
    const int NS = p.num_splits;
    for (int idx = tid; idx < n_heads * D; idx += nthr) {
      const int g = idx / D, d = idx % D;
      float ms[kPushSplits], ls[kPushSplits];
      ms[0] = cl[g * 2];
      ls[0] = cl[g * 2 + 1];
      float M = ms[0];
#pragma unroll
      for (int sp = 1; sp < kPushSplits; ++sp) {
        ms[sp] = sp < NS ? *recv_at(sp, g, D) : -INFINITY;
        ls[sp] = sp < NS ? *recv_at(sp, g, D + 1) : 0.f;
        M = fmaxf(M, ms[sp]);
      }
      float L = 0.f, O = 0.f;
#pragma unroll
      for (int sp = 0; sp < kPushSplits; ++sp) {
        if (sp < NS && ms[sp] > -INFINITY) {
          const float w = fast_exp2(ms[sp] - M);
          L = fmaf(ls[sp], w, L);
          O = fmaf(sp == 0 ? part_o[g * D + d] : *recv_at(sp, g, d), w, O);
        }
      }
      store_out<T>(p, ob + (int64_t)(first_head + g) * p.os[1] + d * p.os[3], O / L);
      if (d == 0) store_ml(p, b, first_head + g, M, L);
      if (p.dead && d == 0) p.dead[(int64_t)b * p.Hq + first_head + g] = L > 0.f ? 0 : 1;
    }
    trace_mark(p, 6);
    return;
}

template <typename T>
__device__ OMX_NI_COMBINE void combine_pull(const DecodeParams& p, const MergeArgs& a) {
  const float* mo = a.mo;
  const int n_ent = a.n_ent, rows = a.rows, first_head = a.first_head, n_heads = a.n_heads, b = a.b, pair = a.pair,
            split = a.split, tid = a.tid, nthr = a.nthr;
  int* s_ticket = a.s_ticket;
  float *cl = a.cl, *recv = a.recv;
  const int D = p.D;
  const int64_t ob = b * p.os[0];
  float* part_o = const_cast<float*>(mo);
  auto recv_at = [&](int sp, int g, int d) { return recv + ((sp - 1) * n_heads + g) * (D + 2) + d; };
  (void)n_ent; (void)rows; (void)pair; (void)split; (void)s_ticket; (void)cl; (void)recv; (void)ob; (void)part_o; (void)recv_at;
    // ---- cluster combine.  Every CTA of the cluster publishes its partial in its own shared memory,
    // one cluster barrier later each CTA pulls all (m, l) pairs (NS x heads x 8 B) and its slice of the
    // output columns from its peers (~215-cycle DSMEM loads), and stores that slice.  Replaces partial
    // write -> fence -> ticket -> L2 read-back (~5 us on the per-CTA timelines) by two cluster barriers.
    const int NS = p.num_splits;
    float* sm_w = cl + 2 * rows;                        // [NS][rows]
    float* sm_l = sm_w + kMaxClusterSplits * rows;      // [NS][rows]
    cluster_arrive_release();
    cluster_wait_acquire();
    trace_mark(p, 4);
    for (int idx = tid; idx < NS * n_heads; idx += nthr) {
      const int sp = idx / n_heads, g = idx % n_heads;
      const float2 ml = ld_dsmem_f2(dsmem_addr(cl + g * 2, sp));
      sm_w[sp * rows + g] = ml.x;
      sm_l[sp * rows + g] = ml.y;
    }
    __syncthreads();
    for (int g = tid >> 5; g < n_heads; g += nthr >> 5) {  // one warp per head, lanes over the splits
      const int ln = tid & 31;
      const float m_a = ln < NS ? sm_w[ln * rows + g] : -INFINITY;
      const float M = warp_max(m_a);
      const float s_a = m_a > -INFINITY ? fast_exp2(m_a - M) : 0.f;
      const float L = warp_sum(ln < NS ? sm_l[ln * rows + g] * s_a : 0.f);
      if (ln < NS) sm_w[ln * rows + g] = s_a * (1.0f / L);
      if (ln == 0 && split == 0) store_ml(p, b, first_head + g, M, L);
      if (p.dead && ln == 0 && split == 0) p.dead[(int64_t)b * p.Hq + first_head + g] = L > 0.f ? 0 : 1;
    }
    __syncthreads();
    trace_mark(p, 5);
    const int D4 = D >> 2;
    const int C = n_heads * D4;
    const int slice = (C + NS - 1) / NS;
    const int c_begin = split * slice, c_end = min(C, c_begin + slice);  // split == rank within the cluster
    for (int col = c_begin + tid; col < c_end; col += nthr) {
      const int g = col / D4, d = (col % D4) * 4;
      const float* src = part_o + g * D + d;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int sp = 0; sp < NS; ++sp) {
        const float4 v = ld_dsmem_f4(dsmem_addr(src, sp));
        const float w = sm_w[sp * rows + g];
        acc.x = fmaf(v.x, w, acc.x); acc.y = fmaf(v.y, w, acc.y);
        acc.z = fmaf(v.z, w, acc.z); acc.w = fmaf(v.w, w, acc.w);
      }
      const int64_t o = ob + (int64_t)(first_head + g) * p.os[1];
      store_out<T>(p, o + (d + 0) * p.os[3], acc.x);
      store_out<T>(p, o + (d + 1) * p.os[3], acc.y);
      store_out<T>(p, o + (d + 2) * p.os[3], acc.z);
      store_out<T>(p, o + (d + 3) * p.os[3], acc.w);
    }
    trace_mark(p, 11);
    // nobody may leave while a peer can still read its shared memory
    cluster_arrive_release();
    cluster_wait_acquire();
    trace_mark(p, 6);
    return;
}

template <typename T, int PF>
__device__ OMX_NI_COMBINE void combine_ticket(const DecodeParams& p, const MergeArgs& a) {
  const float* mo = a.mo;
  const int n_ent = a.n_ent, rows = a.rows, first_head = a.first_head, n_heads = a.n_heads, b = a.b, pair = a.pair,
            split = a.split, tid = a.tid, nthr = a.nthr;
  int* s_ticket = a.s_ticket;
  float *cl = a.cl, *recv = a.recv;
  const int D = p.D;
  const int64_t ob = b * p.os[0];
  float* part_o = const_cast<float*>(mo);
  auto recv_at = [&](int sp, int g, int d) { return recv + ((sp - 1) * n_heads + g) * (D + 2) + d; };
  (void)n_ent; (void)rows; (void)pair; (void)split; (void)s_ticket; (void)cl; (void)recv; (void)ob; (void)part_o; (void)recv_at;
  __threadfence();
  __syncthreads();
  trace_mark(p, 4);
  if (tid == 0) *s_ticket = atomicAdd(&p.counters[pair], 1);
  __syncthreads();
  trace_mark(p, 5);
  if (*s_ticket != p.num_splits - 1) return;
  __threadfence();
  trace_mark(p, 8);
  // ---- last CTA of this (batch, kv-head): combine the split partials.  The combine sits at the very end
  // of the launch's critical path, so it is organised as ONE round trip to L2: every thread owns a float4
  // column (head, 4 features) for a subset of the splits and requests its first 8 partial vectors BEFORE
  // the (m, l) pairs are turned into weights; (A) all (split, head) pairs -> shared memory, (B) one thread
  // per head computes 2^(m_s - M) / L, (C) the prefetched vectors are folded in, the column groups are
  // reduced through shared memory.  (Per-CTA timelines: the earlier three dependent phases took ~6 us
  // for 16 splits x 4 heads.)
  float* sm_w = const_cast<float*>(mo);            // [num_splits][rows]  (the merge inputs are dead)
  float* sm_l = sm_w + kMaxSplits * rows;          // [num_splits][rows]
  float4* red = reinterpret_cast<float4*>(sm_l + kMaxSplits * rows);  // [groups][C]
  const int64_t e0p = (int64_t)pair * p.num_splits * n_heads;
  const int D4 = D >> 2;
  const int C = n_heads * D4;  // float4 columns
  // thread groups: as many as fit the CTA, the splits and the scratch left in `mo` (n_ent * rows * D floats)
  int groups = max(1, min(min(nthr / C, p.num_splits), 4));
  while (groups > 1 && 2 * kMaxSplits * rows + groups * C * 4 > n_ent * rows * D) --groups;
  const int grp = tid / C, col0 = tid % C;
  const bool first_pass = C <= nthr && grp < groups;
  float4 pre[PF];  // PF x groups splits are in flight before the weights exist
  if (first_pass) {
    const float* po = p.ws_o + (e0p + col0 / D4) * D + (col0 % D4) * 4;
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int sp = grp + i * groups;
      if (sp < p.num_splits) pre[i] = __ldcg(reinterpret_cast<const float4*>(po + (int64_t)sp * n_heads * D));
    }
  }
  for (int idx = tid; idx < p.num_splits * n_heads; idx += nthr) {
    const int sp = idx / n_heads, g = idx % n_heads;
    const float2 ml = __ldcg(reinterpret_cast<const float2*>(&p.ws_ml[(e0p + idx) * 2]));
    sm_w[sp * rows + g] = ml.x;
    sm_l[sp * rows + g] = ml.y;
  }
  __syncthreads();
  trace_mark(p, 9);
  // (B) one warp per head, lanes over the splits (kMaxSplits = 64 = 2 per lane), butterfly reductions
  for (int g = tid >> 5; g < n_heads; g += nthr >> 5) {
    const int ln = tid & 31;
    const float m_a = ln < p.num_splits ? sm_w[ln * rows + g] : -INFINITY;
    const float m_b = ln + 32 < p.num_splits ? sm_w[(ln + 32) * rows + g] : -INFINITY;
    const float M = warp_max(fmaxf(m_a, m_b));
    const float s_a = m_a > -INFINITY ? fast_exp2(m_a - M) : 0.f;
    const float s_b = m_b > -INFINITY ? fast_exp2(m_b - M) : 0.f;
    float L = ln < p.num_splits ? sm_l[ln * rows + g] * s_a : 0.f;
    if (ln + 32 < p.num_splits) L = fmaf(sm_l[(ln + 32) * rows + g], s_b, L);
    L = warp_sum(L);
    const float inv = 1.0f / L;
    if (ln < p.num_splits) sm_w[ln * rows + g] = s_a * inv;
    if (ln + 32 < p.num_splits) sm_w[(ln + 32) * rows + g] = s_b * inv;
    if (ln == 0) store_ml(p, b, first_head + g, M, L);
    if (p.dead && ln == 0) p.dead[(int64_t)b * p.Hq + first_head + g] = L > 0.f ? 0 : 1;
  }
  __syncthreads();
  trace_mark(p, 10);
  auto fold = [](float4& a, const float4& v, float w) {
    a.x = fmaf(v.x, w, a.x); a.y = fmaf(v.y, w, a.y); a.z = fmaf(v.z, w, a.z); a.w = fmaf(v.w, w, a.w);
  };
  auto emit = [&](int col, const float4& a) {
    const int g = col / D4, d = (col % D4) * 4;
    const int64_t o = ob + (int64_t)(first_head + g) * p.os[1];
    store_out<T>(p, o + (d + 0) * p.os[3], a.x);
    store_out<T>(p, o + (d + 1) * p.os[3], a.y);
    store_out<T>(p, o + (d + 2) * p.os[3], a.z);
    store_out<T>(p, o + (d + 3) * p.os[3], a.w);
  };
  if (C <= nthr) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (first_pass) {
      const int g = col0 / D4;
      const float* po = p.ws_o + (e0p + g) * D + (col0 % D4) * 4;
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        const int sp = grp + i * groups;
        if (sp < p.num_splits) fold(acc, pre[i], sm_w[sp * rows + g]);
      }
      // more splits than one batch of PF (a single (batch, kv-head) pair spread over 64 CTAs, e.g. one rank of
      // the kv-head-sharded C5): further batches of PF loads in flight, not PF/4 -- 64 splits = 4 round trips
      // to L2 instead of 13
      if (p.num_splits <= (PF + 4) * groups) {  // a short tail (C5 on one GPU: 18 splits) keeps the plain loop
#pragma unroll 4
        for (int sp = grp + PF * groups; sp < p.num_splits; sp += groups)
          fold(acc, __ldcg(reinterpret_cast<const float4*>(po + (int64_t)sp * n_heads * D)), sm_w[sp * rows + g]);
      } else
      for (int base = grp + PF * groups; base < p.num_splits; base += PF * groups) {
#pragma unroll
        for (int i = 0; i < PF; ++i) {
          const int sp = base + i * groups;
          if (sp < p.num_splits) pre[i] = __ldcg(reinterpret_cast<const float4*>(po + (int64_t)sp * n_heads * D));
        }
#pragma unroll
        for (int i = 0; i < PF; ++i) {
          const int sp = base + i * groups;
          if (sp < p.num_splits) fold(acc, pre[i], sm_w[sp * rows + g]);
        }
      }
    }
    if (groups > 1) {
      if (first_pass) red[grp * C + col0] = acc;
      __syncthreads();
      if (tid < C) {
        for (int gi = 1; gi < groups; ++gi) {
          const float4 v = red[gi * C + tid];
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        emit(tid, acc);
      }
    } else if (tid < C) {
      emit(tid, acc);
    }
  } else {  // more columns than threads (G = 16 at D = 128, D = 256): each thread walks its columns
    for (int col = tid; col < C; col += nthr) {
      const int g = col / D4;
      const float* po = p.ws_o + (e0p + g) * D + (col % D4) * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int sp = 0; sp < p.num_splits; ++sp)
        fold(acc, __ldcg(reinterpret_cast<const float4*>(po + (int64_t)sp * n_heads * D)), sm_w[sp * rows + g]);
      emit(col, acc);
    }
  }
  trace_mark(p, 11);
  if (tid == 0) p.counters[pair] = 0;  // self-reset for the next launch
  peer_signal(p, tid);
  trace_mark(p, 6);
}

// Merge per-warp states -> out (single split) or workspace + last-CTA combine.
// mo: [n_ent][rows][D] floats, mml: [n_ent][rows][2]; rows = row pitch of the entries.
template <typename T, int PF>
__device__ __forceinline__ void merge_and_store(const DecodeParams& p, const float* mo, const float* mml,
                                                int n_ent, int rows, bool has_nt, const float* nt_m,
                                                const float* nt_v, int first_head, int n_heads, int b,
                                                int pair, int split, int tid, int nthr, int* s_ticket,
                                                float* cl /* cluster scratch, cluster_scratch_floats(rows) */,
                                                float* recv = nullptr /* push-combine receive slots */,
                                                unsigned gs_tag = 0 /* this launch's tag (gsync == 2) */) {
  const int D = p.D;
  const int64_t ob = b * p.os[0];
  const bool use_cluster = p.cluster && p.num_splits > 1;
  const bool push = use_cluster && p.push_combine && recv != nullptr;
  // receive slot of split s >= 1 inside RANK 0's shared memory: [s - 1][head][D | m | l]
  auto recv_at = [&](int sp, int g, int d) { return recv + ((sp - 1) * n_heads + g) * (D + 2) + d; };
  float* part_o = const_cast<float*>(mo);  // [n_heads][D]: written in place over warp 0's block (same owner)
  if (push) cluster_wait_plain();  // pairs with the arrive at kernel entry: rank 0 is running, its slots exist
  const int dsh = 31 - __clz(D);  // head dims are powers of two (decode_supported)
  // (one element per thread: a float4-column variant halved the active threads and measured slower)
  for (int idx = tid; idx < n_heads * D; idx += nthr) {
    const int g = idx >> dsh, d = idx & (D - 1);
    float M = has_nt ? nt_m[g] : -INFINITY;
    for (int w = 0; w < n_ent; ++w) M = fmaxf(M, mml[(w * rows + g) * 2]);
    float L = 0.f, O = 0.f;
    for (int w = 0; w < n_ent; ++w) {
      const float mw = mml[(w * rows + g) * 2];
      if (mw > -INFINITY) {
        const float sc = fast_exp2(mw - M);
        L = fmaf(mml[(w * rows + g) * 2 + 1], sc, L);
        O = fmaf(mo[(w * rows + g) * D + d], sc, O);
      }
    }
    if (has_nt) {
      const float sc = fast_exp2(nt_m[g] - M);
      L += sc;
      O = fmaf(nt_v[d], sc, O);
    }
    if (p.num_splits == 1) {
      store_out<T>(p, ob + (int64_t)(first_head + g) * p.os[1] + d * p.os[3], O / L);
      if (d == 0) store_ml(p, b, first_head + g, M, L);
      if (p.dead && d == 0) p.dead[(int64_t)b * p.Hq + first_head + g] = L > 0.f ? 0 : 1;
    } else if (push && split != 0) {
      st_dsmem_f32(dsmem_addr(recv_at(split, g, d), 0), O);
      if (d == 0) {
        st_dsmem_f32(dsmem_addr(recv_at(split, g, D), 0), M);
        st_dsmem_f32(dsmem_addr(recv_at(split, g, D + 1), 0), L);
      }
    } else if (use_cluster) {
      part_o[g * D + d] = O;  // == mo[(0 * rows + g) * D + d], read above by this thread only
      if (d == 0) {
        cl[g * 2] = M;
        cl[g * 2 + 1] = L;
      }
    } else if (p.gsync == 2) {
      uint2* w = p.ws_w + (((int64_t)pair * p.num_splits + split) * n_heads + g) * (D + 2);
      dd::st_ll(reinterpret_cast<unsigned long long*>(w + d), __float_as_uint(O), gs_tag);
      if (d == 0) {
        dd::st_ll(reinterpret_cast<unsigned long long*>(w + D), __float_as_uint(M), gs_tag);
        dd::st_ll(reinterpret_cast<unsigned long long*>(w + D + 1), __float_as_uint(L), gs_tag);
      }
    } else {
      const int64_t e = ((int64_t)pair * p.num_splits + split) * n_heads + g;
      p.ws_o[e * D + d] = O;
      if (d == 0) {
        p.ws_ml[e * 2] = M;
        p.ws_ml[e * 2 + 1] = L;
      }
    }
  }
  if (p.num_splits == 1) {
    peer_signal(p, tid);
    trace_mark(p, 6);
    return;
  }
  trace_mark(p, 3);
  const MergeArgs ma{mo, n_ent, rows, first_head, n_heads, b, pair, split, tid, nthr, s_ticket, cl, recv};
  if (push) {
    combine_push<T>(p, ma);
  } else if (use_cluster) {
    combine_pull<T>(p, ma);
  } else if (p.gsync) {
    gsync_combine<T>(p, const_cast<float*>(mo), first_head, n_heads, b, pair, split, tid, nthr, gs_tag);
  } else {
    combine_ticket<T, PF>(p, ma);
  }
}

// ============================================================ 16-bit, D = 128: TMA + mma.sync
// NSTAGE consumer warps + 1 producer warp.  Consumer warp w owns stage w and its full/empty
// mbarrier pair (tile t -> warp t % NSTAGE -> stage t % NSTAGE): every barrier then has exactly
// one waiter that consumes its phases in order, which the parity protocol requires (a waiter that
// skipped ahead to a later phase of a fresh barrier would fall straight through).
template <typename T, int NSTAGE, bool HI, int MINB>
__global__ void __launch_bounds__((NSTAGE + 1) * 32, MINB)
decode_hmma_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ DecodeParams p) {
  constexpr int D = 128;
  constexpr int NW = NSTAGE;
  constexpr int NTHR = (NW + 1) * 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  T* q_s = reinterpret_cast<T*>(smem + NSTAGE * kStageBytes);  // [16][kQPitch]
  float* nt_k = reinterpret_cast<float*>(q_s + 16 * kQPitch);
  float* nt_v = nt_k + D;
  float* nt_m = nt_v + D;
  __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE];
  __shared__ int s_ticket;
  __shared__ PrologueSmem s_pro;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  const int G = p.G;
  const int pair = b * p.Hkv + hk;
  if (p.push_combine) cluster_arrive_relaxed();  // (see cluster_wait_plain in merge_and_store)
  // ---- before the dependency wait: shared memory, and cache rows no running kernel can still be writing
  int tile_begin = split * p.tiles_per_split, my_tiles = 0;
  const int* bt_row = nullptr;
  uint64_t pol = 0;
  int pf_next = 0;  // (producer lane) next tile of this CTA to request into L2
  auto l2_upto = [&](int from, int hi) {
#pragma unroll 1
    for (pf_next = max(pf_next, from); pf_next < hi; ++pf_next) {
      const int k0 = p.paged ? 0 : (tile_begin + pf_next) * kTile;
      const int c = p.paged ? __ldg(bt_row + tile_begin + pf_next) : b;
      tma_prefetch_l2_4d(&tmK, 0, k0, hk, c);
      tma_prefetch_l2_4d(&tmK, 64, k0, hk, c);
      tma_prefetch_l2_4d(&tmV, 0, k0, hk, c);
      tma_prefetch_l2_4d(&tmV, 64, k0, hk, c);
    }
  };
  auto issue = [&](int t, int lim) {
    const int st = t % NSTAGE;
    uint8_t* sb = stages + st * kStageBytes;
    // contiguous cache: rows [key0, key0 + 64) of (b, hk); paged: the whole page block_table[b][tile]
    const int key0 = p.paged ? 0 : (tile_begin + t) * kTile;
    const int c3 = p.paged ? __ldg(bt_row + tile_begin + t) : b;
    mbar_expect_tx(&full_bar[st], kStageBytes);
    tma_load_4d(sb, &tmK, &full_bar[st], 0, key0, hk, c3, pol);
    tma_load_4d(sb + kBoxBytes, &tmK, &full_bar[st], 64, key0, hk, c3, pol);
    tma_load_4d(sb + 2 * kBoxBytes, &tmV, &full_bar[st], 0, key0, hk, c3, pol);
    tma_load_4d(sb + 3 * kBoxBytes, &tmV, &full_bar[st], 64, key0, hk, c3, pol);
    if (p.l2_ahead > 0) l2_upto(t + 1, min(lim, t + 1 + p.l2_ahead));
  };
  int n_early = 0;  // (producer lane) tiles requested before the wait
  if (tid == NW * 32) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    pol = policy_evict_first();
    if (p.stable_rows > 0) {  // host: static position, contiguous cache
      const int mine = max(0, min(p.tiles_per_split, (p.n_mem + kTile - 1) / kTile - tile_begin));
      const int lim = max(0, min(mine, p.stable_rows / kTile - tile_begin));
      n_early = min(lim, NSTAGE);
      for (int t = 0; t < n_early; ++t) issue(t, lim);
      l2_upto(n_early, min(lim, n_early + p.l2_early));
    }
  }
  pdl_wait_prior_grid();
  pdl_release_next_grid();
  trace_mark(p, 0);
  const unsigned gs_tag = p.gsync == 2 ? __ldcg(p.gs_seq) + 1u : 0u;  // (needed at the merge: in flight under the loop)
  if (p.paged && __ldg(p.pos_dev + b) < 0) {  // inactive slot (whole CTA, whole cluster: same b)
    gs_leave_unused(p, gs_tag, tid);
    return;
  }
  const DecodeDyn dy = load_dyn(p, b);
  const int n_tiles = (dy.n_mem + kTile - 1) / kTile;
  tile_begin = split * dy.tps;
  my_tiles = max(0, min(dy.tps, n_tiles - tile_begin));
  const bool has_nt = p.fused && p.append && split == p.num_splits - 1;
  bt_row = p.paged ? p.block_table + (size_t)b * p.bt_stride : nullptr;
  if (p.paged && has_nt && hk == 0 && tid == 0) p.lens_out[b] = dy.Lk;  // the sequence's length after this step

  // The producer lane puts the first NSTAGE tiles in flight before (or, on one-wave grids, while) the CTA
  // stages q: the K/V stream does not depend on q, and with only a dozen tiles per CTA (single sequence, many
  // splits) the q round trip would otherwise sit in front of the whole pipeline.
  const int first = min(my_tiles, NSTAGE);
  // One-wave grids (MINB == 1, single-sequence decode): the producer warp and the consumer warps run apart --
  // the producer lane puts the first tiles in flight (~1 us of serial issue)
  // WHILE the consumer warps stage q; barrier 2 = the consumer warps among themselves, barrier 1 = "q is
  // staged" for the producer warp (consumers arrive without blocking).  A/B in one box: fused step -1..-2.6 %.
  // Multi-wave grids (C2) keep the single-barrier flow, which measured 0.3 % faster there.
  constexpr bool kDecouple = MINB == 1;
  if constexpr (kDecouple) __syncthreads();  // barriers initialised before anyone waits on them
  if (tid == NW * 32) {
    for (int t = n_early; t < first; ++t) issue(t, my_tiles);
    if (p.trace) {
      unsigned long long tt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
      p.trace[(size_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) * 16 + 7] = tt;
    }
  }
  if constexpr (kDecouple) {
    if (warp == NW) {
      __syncwarp();  // (the producer warp meets barrier 1, "q is staged", after its issue loop)
    } else {
      stage_q<T>(p, q_s, kQPitch, 16, hk * G, G, b, hk, tid, NW * 32, s_pro, nt_k, nt_v, has_nt, dy, [] {},
                 [] { named_bar_sync(2, NW * 32); });
      named_bar_sync(2, NW * 32);
      named_bar_arrive(1, NTHR);
      trace_mark(p, 1);
    }
  } else {
    stage_q<T>(p, q_s, kQPitch, 16, hk * G, G, b, hk, tid, NTHR, s_pro, nt_k, nt_v, has_nt, dy, [] {},
               [] { __syncthreads(); });
    __syncthreads();
    trace_mark(p, 1);
  }

  // consumer state (declared at function scope so the merge below runs after CTA-wide barriers
  // that every warp reaches at the same program point).
  //
  // Operand roles are SWAPPED relative to the textbook S = Q K^T: the keys (resp. features of V) take the
  // 16-row M side of mma.m16n8k16 and the G <= 8 query heads the 8-wide N side,
  //     S^T[key, head] = K[key, :] . Q[head, :]        O^T[feat, head] += V^T[feat, key] . P^T[key, head]
  // so no MMA row is zero padding: 64 MMAs per 64-key tile instead of 128 (G = 4..8).  With one warp per
  // tile the legacy-MMA issue rate, not HBM, bounded every grid below ~2 tiles in flight per scheduler
  // (B = 8, ctx 4096: 4.3 TB/s).  P^T reaches its B-fragment layout through movmatrix (8x8 b16
  // transpose).  G > 8 (HI): a second N block (heads 8..15) reuses the same A fragments.
  constexpr int NG = HI ? 2 : 1;  // head groups of 8
  const int kr = lane >> 2;       // C-fragment row (key resp. feature within 8)
  const int hc = (lane & 3) * 2;  // C-fragment columns hc, hc + 1 (head within the group)
  float ot[NG][8][4];             // O^T: [group][16-feature block][frag]
#pragma unroll
  for (int n = 0; n < NG; ++n)
#pragma unroll
    for (int i = 0; i < 8; ++i) ot[n][i][0] = ot[n][i][1] = ot[n][i][2] = ot[n][i][3] = 0.f;
  float mrun[NG][2], lrun[NG][2];  // running max / partial row sum of heads hc, hc + 1 of each group
#pragma unroll
  for (int n = 0; n < NG; ++n) {
    mrun[n][0] = mrun[n][1] = -INFINITY;
    lrun[n][0] = lrun[n][1] = 0.f;
  }

  if (warp == NW) {
    // ------------------------------------------------ producer warp (first tiles already in flight)
    // The K/V stream comes first: the new token (norm + rope of k_new, the cache row store, its score) is only
    // needed by the merge, so it runs AFTER the last tile has been issued, under the consumers' last tiles.
    // (Doing it first kept the appending CTAs' rings from refilling for ~1.5 us: those CTAs -- one per (batch,
    // kv-head) -- finished ~2 us after the rest and set the launch's end; per-CTA timelines, B = 8 per GPU.)
    if (lane == 0) {
      for (int t = first; t < my_tiles; ++t) {
        mbar_wait(&empty_bar[t % NSTAGE], ((t / NSTAGE) - 1) & 1);
        issue(t, my_tiles);
      }
    }
    __syncwarp();
    if constexpr (kDecouple) named_bar_sync(1, NTHR);  // q is staged (the consumers arrived long ago)
    if (has_nt) new_token<T>(p, q_s, kQPitch, G, b, hk, lane, nt_k, nt_v, nt_m, s_pro, dy);

  } else {
    // ------------------------------------------------ consumer warps
    using MMA = Mma16816<T>;
    // Q as the B operand: b0 = Q[head kr][16 ks + hc, +1], b1 = the same 8 features further
    uint32_t qb[NG][8][2];
#pragma unroll
    for (int n = 0; n < NG; ++n)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        qb[n][ks][0] = *reinterpret_cast<const uint32_t*>(&q_s[(n * 8 + kr) * kQPitch + ks * 16 + hc]);
        qb[n][ks][1] = *reinterpret_cast<const uint32_t*>(&q_s[(n * 8 + kr) * kQPitch + ks * 16 + hc + 8]);
      }
    const int r8 = lane & 7, mi = lane >> 3;

    for (int t = warp; t < my_tiles; t += NW) {
      const int st = t % NSTAGE;
      mbar_wait(&full_bar[st], (t / NSTAGE) & 1);
      const uint32_t sbK = smem_u32(stages + st * kStageBytes);
      const uint32_t sbV = sbK + 2 * kBoxBytes;

      // S^T = K Q^T  (64 keys x 8 heads per group): A = K rows through ldmatrix (a0: keys 0-7 / feats 0-7,
      // a1: keys 8-15, a2: feats 8-15, a3: both)
      float sa[NG][4][4];
#pragma unroll
      for (int n = 0; n < NG; ++n)
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) sa[n][kb][0] = sa[n][kb][1] = sa[n][kb][2] = sa[n][kb][3] = 0.f;
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint32_t rowaddr = sbK + (uint32_t)(kb * 16 + (mi & 1) * 8 + r8) * 128u;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const int c = ks * 2 + (mi >> 1);  // 16-byte chunk of the 256-byte key row
          uint32_t af[4];
          ldmatrix_x4(af[0], af[1], af[2], af[3],
                      rowaddr + (uint32_t)(c >> 3) * kBoxBytes + (uint32_t)(((c & 7) ^ r8) << 4));
#pragma unroll
          for (int n = 0; n < NG; ++n) MMA::run(sa[n][kb], af, qb[n][ks][0], qb[n][ks][1]);
        }
      }

      // online softmax (log2 domain): this thread holds, for heads hc / hc + 1 of each group, the keys
      // kb * 16 + kr (frag 0, 1) and + 8 (frag 2, 3)
      const int key_base = (tile_begin + t) * kTile;
      const bool partial = key_base + kTile > dy.n_mem;
      uint32_t pb[NG][4][2];  // P^T as the B operand of O^T += V^T P^T
#pragma unroll
      for (int n = 0; n < NG; ++n) {
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = key_base + kb * 16 + kr + (e >> 1) * 8;
            const int head = n * 8 + hc + (e & 1);
            const bool dead = partial && (key >= dy.n_mem);
            float v = dead ? -INFINITY : sa[n][kb][e] * p.scale_log2;
            if (p.mask_kind && !dead && head < G) v = mask_score<T>(p, v, b, hk * G + head, key);  // launch-uniform
            sa[n][kb][e] = v;
            mx[e & 1] = fmaxf(mx[e & 1], v);
          }
        float cf[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float m = mx[h];
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
          const float mo_ = mrun[n][h];
          mrun[n][h] = fmaxf(mo_, m);
          // (a masked row can have seen no key yet: subtract 0 instead of -inf so that 2^(-inf - m) = 0)
          mx[h] = mrun[n][h] == -INFINITY ? 0.f : mrun[n][h];
          cf[h] = fast_exp2(mo_ - mx[h]);
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const float p0 = fast_exp2(sa[n][kb][0] - mx[0]), p1 = fast_exp2(sa[n][kb][1] - mx[1]);
          const float p2 = fast_exp2(sa[n][kb][2] - mx[0]), p3 = fast_exp2(sa[n][kb][3] - mx[1]);
          rs[0] += p0 + p2;
          rs[1] += p1 + p3;
          // [key kr][heads hc, hc+1] -> transpose -> [head kr][keys hc, hc+1] = the B fragment
          pb[n][kb][0] = movmatrix_trans(MMA::pack(p0, p1));
          pb[n][kb][1] = movmatrix_trans(MMA::pack(p2, p3));
        }
        lrun[n][0] = lrun[n][0] * cf[0] + rs[0];
        lrun[n][1] = lrun[n][1] * cf[1] + rs[1];
#pragma unroll
        for (int fb = 0; fb < 8; ++fb) {
          ot[n][fb][0] *= cf[0];
          ot[n][fb][1] *= cf[1];
          ot[n][fb][2] *= cf[0];
          ot[n][fb][3] *= cf[1];
        }
      }

      // O^T += V^T P^T  (128 features x 8 heads per group): A = V^T through ldmatrix.trans (a0: feats 0-7 /
      // keys 0-7, a1: feats 8-15, a2: keys 8-15, a3: both)
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint32_t rowaddr = sbV + (uint32_t)(kb * 16 + (mi >> 1) * 8 + r8) * 128u;
#pragma unroll
        for (int fb = 0; fb < 8; ++fb) {
          const int c = fb * 2 + (mi & 1);
          uint32_t af[4];
          ldmatrix_x4_trans(af[0], af[1], af[2], af[3],
                            rowaddr + (uint32_t)(c >> 3) * kBoxBytes + (uint32_t)(((c & 7) ^ r8) << 4));
#pragma unroll
          for (int n = 0; n < NG; ++n) MMA::run(ot[n][fb], af, pb[n][kb][0], pb[n][kb][1]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
    }
#pragma unroll
    for (int n = 0; n < NG; ++n)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float l = lrun[n][h];
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
        lrun[n][h] = l;
      }
  }
  __syncthreads();  // every stage consumed -> the ring is reused for the merge
  trace_mark(p, 2);
  if (warp < NW) {
    float* mo = reinterpret_cast<float*>(stages);  // [NW][16][128]
    float* mml = mo + NW * 16 * D;                 // [NW][16][2]
#pragma unroll
    for (int n = 0; n < NG; ++n)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int head = n * 8 + hc + h;
        if (head < G) {
#pragma unroll
          for (int fb = 0; fb < 8; ++fb) {
            mo[(warp * 16 + head) * D + fb * 16 + kr] = ot[n][fb][h];
            mo[(warp * 16 + head) * D + fb * 16 + kr + 8] = ot[n][fb][2 + h];
          }
          if (kr == 0) {
            mml[(warp * 16 + head) * 2] = mrun[n][h];
            mml[(warp * 16 + head) * 2 + 1] = lrun[n][h];
          }
        }
      }
  }
  __syncthreads();  // merge inputs visible
  const float* mo = reinterpret_cast<const float*>(stages);
  merge_and_store<T, 16>(p, mo, mo + NW * 16 * D, NW, 16, has_nt, nt_m, nt_v, hk * G, G, b, pair, split, tid,
                         NTHR, &s_ticket, const_cast<float*>(mo) + NW * 16 * (D + 2),
                         p.push_combine ? nt_m + 16 : nullptr, gs_tag);
}

// ============================================================ generic: CUDA cores
// D = Dv = 32*VE, GT query heads per CTA, 8 warps; lane owns features [lane*VE, lane*VE+VE).
template <typename T, int VE>
__device__ __forceinline__ void load_row(const T* p, float (&f)[VE]) {
  if constexpr (sizeof(T) * VE == 16) {
    union { uint4 r; T t[VE]; } u;
    u.r = *reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int e = 0; e < VE; ++e) f[e] = Num<T>::to_f(u.t[e]);
  } else if constexpr (sizeof(T) * VE == 8) {
    union { uint2 r; T t[VE]; } u;
    u.r = *reinterpret_cast<const uint2*>(p);
#pragma unroll
    for (int e = 0; e < VE; ++e) f[e] = Num<T>::to_f(u.t[e]);
  } else if constexpr (sizeof(T) * VE == 32) {
    union { uint4 r[2]; T t[VE]; } u;
    u.r[0] = reinterpret_cast<const uint4*>(p)[0];
    u.r[1] = reinterpret_cast<const uint4*>(p)[1];
#pragma unroll
    for (int e = 0; e < VE; ++e) f[e] = Num<T>::to_f(u.t[e]);
  } else if constexpr (sizeof(T) * VE == 4) {
    union { uint32_t r; T t[VE]; } u;
    u.r = *reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int e = 0; e < VE; ++e) f[e] = Num<T>::to_f(u.t[e]);
  } else {
#pragma unroll
    for (int e = 0; e < VE; ++e) f[e] = Num<T>::to_f(p[e]);
  }
}

// VE elements of one K/V row as loaded (128-bit pieces where the size allows), converted on use
template <typename T, int VE>
struct RawRow {
  static constexpr int BYTES = (int)sizeof(T) * VE;
  union U {
    uint4 q[BYTES >= 16 ? BYTES / 16 : 1];
    uint2 d;
    uint32_t w;
    T t[VE];
    __device__ U() {}
  } u;
  __device__ __forceinline__ void load(const T* p) {
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
      for (int i = 0; i < BYTES / 16; ++i) u.q[i] = reinterpret_cast<const uint4*>(p)[i];
    } else if constexpr (BYTES == 8) {
      u.d = *reinterpret_cast<const uint2*>(p);
    } else if constexpr (BYTES == 4) {
      u.w = *reinterpret_cast<const uint32_t*>(p);
    } else {
#pragma unroll
      for (int e = 0; e < VE; ++e) u.t[e] = p[e];
    }
  }
  __device__ __forceinline__ float f(int e) const { return Num<T>::to_f(u.t[e]); }
};

constexpr int kSimtWarps = 8;
// KPW = keys in flight per warp iteration: 4 when many CTAs share an SM (throughput regime), 16 for
// one-wave grids (single-sequence decode), where a CTA's whole 128-key share is then requested up front
// instead of through four dependent round trips to HBM.
// ST (staged; one-wave grids, contiguous rows): ONE thread asks for the CTA's whole key range with two bulk copies
// into shared memory -- all of it in flight from the first microsecond, none of it in registers -- and the key
// loop is the small 4-key body run from shared memory (the 16-key body is ~1000 straight-line instructions per
// head pair: the per-CTA timelines show its first and only walk taking 5.4 us on a cold instruction cache).
template <typename T, int VE, int GT, int KPW, bool ST = false>
__global__ void __launch_bounds__(kSimtWarps * 32)
decode_simt_kernel(const __grid_constant__ DecodeParams p) {
  constexpr int kSimtKeys = KPW;
  constexpr int D = 32 * VE;
  constexpr int NTHR = kSimtWarps * 32;
  extern __shared__ uint8_t smem_raw[];
  T* q_s = reinterpret_cast<T*>(smem_raw);         // [GT][D] (16-byte aligned rows)
  float* mo = reinterpret_cast<float*>(q_s + GT * D);  // [warps][GT][D]
  float* mml = mo + kSimtWarps * GT * D;           // [warps][GT][2]
  float* nt_k = mml + kSimtWarps * GT * 2;         // [D]
  float* nt_v = nt_k + D;
  float* nt_m = nt_v + D;                          // [GT]
  __shared__ int s_ticket;
  __shared__ PrologueSmem s_pro;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, b = blockIdx.z;
  const int groups = p.G / GT;
  const int hk = blockIdx.y / groups, gsub = blockIdx.y % groups;
  const int first_head = hk * p.G + gsub * GT;
  const int pair = (b * p.Hkv + hk) * groups + gsub;
  __shared__ uint64_t st_bar;
  T* ks_s = reinterpret_cast<T*>(smem_raw + p.stage_off);
  T* vs_s = ks_s + (size_t)p.stage_keys * D;
  bool st_issued = false;  // (thread 0)
  auto stage_issue = [&](int k0, int k1) {  // rows [k0, k1) of (b, hk): contiguous in the cache (host check)
    const uint32_t bytes = (uint32_t)(k1 - k0) * D * (uint32_t)sizeof(T);
    mbar_expect_tx(&st_bar, 2 * bytes);
    bulk_load_1d(ks_s, (const T*)p.k + b * p.ks[0] + hk * p.ks[1] + (int64_t)k0 * p.ks[2], bytes, &st_bar);
    bulk_load_1d(vs_s, (const T*)p.v + b * p.vs[0] + hk * p.vs[1] + (int64_t)k0 * p.vs[2], bytes, &st_bar);
    st_issued = true;
  };
  if constexpr (ST) {
    if (tid == 0) {
      mbar_init(&st_bar, 1);
      mbar_fence_init();
      if (p.stable_rows > 0) {  // rows older than the stream's previous kernel: requested before the dependency wait
        const int k0 = split * p.tiles_per_split * kTile, k1 = min(p.n_mem, k0 + p.tiles_per_split * kTile);
        if (k1 > k0 && k1 <= p.stable_rows) stage_issue(k0, k1);
      }
    }
  }
  pdl_wait_prior_grid();
  trace_mark(p, 0);
  const unsigned gs_tag = p.gsync == 2 ? __ldcg(p.gs_seq) + 1u : 0u;
  if (p.paged && __ldg(p.pos_dev + b) < 0) {  // inactive slot
    gs_leave_unused(p, gs_tag, tid);
    return;
  }
  const DecodeDyn dy = load_dyn(p, b);
  const int keys_per_split = dy.tps * kTile;
  const int kbeg = split * keys_per_split;
  const int kend = min(dy.n_mem, kbeg + keys_per_split);
  if constexpr (ST) {
    if (tid == 0 && !st_issued && kend > kbeg) stage_issue(kbeg, kend);
  }
  const bool has_nt = p.fused && p.append && split == p.num_splits - 1;
  if (p.paged && has_nt && hk == 0 && gsub == 0 && tid == 0) p.lens_out[b] = dy.Lk;

  // K/V do not depend on q: the first batch of rows is requested before the prologue's round trip
  const T* kb = (const T*)p.k + (p.paged ? 0 : b * p.ks[0]) + hk * p.ks[1] + lane * VE;
  const T* vb = (const T*)p.v + (p.paged ? 0 : b * p.vs[0]) + hk * p.vs[1] + lane * VE;
  const int* bt_row = p.paged ? p.block_table + (size_t)b * p.bt_stride : nullptr;
  // rows stay in their storage type until they are used, so that nothing waits on the loads early
  RawRow<T, VE> kraw[kSimtKeys], vraw[kSimtKeys];
  auto load_kv = [&](int j0) {
    if constexpr (ST) {
#pragma unroll
      for (int u = 0; u < kSimtKeys; ++u) kraw[u].load(ks_s + (size_t)(min(j0 + u, kend - 1) - kbeg) * D + lane * VE);
#pragma unroll
      for (int u = 0; u < kSimtKeys; ++u) vraw[u].load(vs_s + (size_t)(min(j0 + u, kend - 1) - kbeg) * D + lane * VE);
      return;
    }
    if (p.paged) {
      // row j = row j % 64 of page block_table[b][j / 64]; j0 is a multiple of KPW and KPW divides 64, so the
      // (clamped) rows of one batch share a page
      const int64_t pg = __ldg(bt_row + (j0 >> 6));
      const T* kp = kb + pg * p.ks[0];
      const T* vp = vb + pg * p.vs[0];
#pragma unroll
      for (int u = 0; u < kSimtKeys; ++u) kraw[u].load(kp + (int64_t)(min(j0 + u, kend - 1) & 63) * p.ks[2]);
#pragma unroll
      for (int u = 0; u < kSimtKeys; ++u) vraw[u].load(vp + (int64_t)(min(j0 + u, kend - 1) & 63) * p.vs[2]);
      return;
    }
#pragma unroll
    for (int u = 0; u < kSimtKeys; ++u) kraw[u].load(kb + (int64_t)min(j0 + u, kend - 1) * p.ks[2]);
#pragma unroll
    for (int u = 0; u < kSimtKeys; ++u) vraw[u].load(vb + (int64_t)min(j0 + u, kend - 1) * p.vs[2]);
  };
  const int j_first = kbeg + warp * kSimtKeys;

  stage_q<T>(p, q_s, D, GT, first_head, GT, b, hk, tid, NTHR, s_pro, nt_k, nt_v, has_nt, dy, [&] {
    if constexpr (!ST) {
      if (j_first < kend) load_kv(j_first);
    }
  }, [] { __syncthreads(); });
  __syncthreads();
  trace_mark(p, 1);

  // the new row is appended once per kv head (gsub == 0 writes it); every group scores it.  (Behind the key loop
  // instead -- only the merge needs it -- measured slower: here it runs while the CTA's rows are still in flight.)
  if (has_nt && warp == kSimtWarps - 1)
    new_token<T>(p, q_s, D, GT, b, hk, lane, nt_k, nt_v, nt_m, s_pro, dy, /*write_cache=*/gsub == 0);

  float qr[GT][VE], acc[GT][VE], m[GT], l[GT];
#pragma unroll
  for (int g = 0; g < GT; ++g) {
    load_row<T, VE>(q_s + g * D + lane * VE, qr[g]);
    m[g] = -INFINITY;
    l[g] = 0.f;
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[g][e] = 0.f;
  }
  if constexpr (ST) {
    if (kend > kbeg) mbar_wait(&st_bar, 0);  // the staged rows have landed (the barrier was initialised before the
  }                                          // CTA-wide syncs of stage_q)
#pragma unroll 1
  for (int j0 = j_first; j0 < kend; j0 += kSimtWarps * kSimtKeys) {
    if (ST || j0 != j_first) load_kv(j0);
#pragma unroll
    for (int g = 0; g < GT; ++g) {
      float s[kSimtKeys];
#pragma unroll
      for (int u = 0; u < kSimtKeys; ++u) {
        float a = 0.f;
#pragma unroll
        for (int e = 0; e < VE; ++e) a = fmaf(qr[g][e], kraw[u].f(e), a);
        s[u] = a;
      }
#pragma unroll
      for (int u = 0; u < kSimtKeys; ++u) s[u] = warp_sum(s[u]);
      float mx = -INFINITY;
#pragma unroll
      for (int u = 0; u < kSimtKeys; ++u) {
        s[u] = (j0 + u < kend) ? s[u] * p.scale_log2 : -INFINITY;
        if (p.mask_kind && j0 + u < kend) s[u] = mask_score<T>(p, s[u], b, first_head + g, j0 + u);
        mx = fmaxf(mx, s[u]);
      }
      const float mo_ = m[g];
      m[g] = fmaxf(m[g], mx);
      const float mn = m[g] == -INFINITY ? 0.f : m[g];  // masked rows may have seen no key yet
      const float c = fast_exp2(mo_ - mn);
      float rs = 0.f;
#pragma unroll
      for (int u = 0; u < kSimtKeys; ++u) {
        s[u] = fast_exp2(s[u] - mn);
        rs += s[u];
      }
      l[g] = l[g] * c + rs;
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        float a = acc[g][e] * c;
#pragma unroll
        for (int u = 0; u < kSimtKeys; ++u) a = fmaf(s[u], vraw[u].f(e), a);
        acc[g][e] = a;
      }
    }
  }
#pragma unroll
  for (int g = 0; g < GT; ++g) {
#pragma unroll
    for (int e = 0; e < VE; ++e) mo[(warp * GT + g) * D + lane * VE + e] = acc[g][e];
    if (lane == 0) {
      mml[(warp * GT + g) * 2] = m[g];
      mml[(warp * GT + g) * 2 + 1] = l[g];
    }
  }
  __syncthreads();
  pdl_release_next_grid();
  trace_mark(p, 2);
  merge_and_store<T, 4>(p, mo, mml, kSimtWarps, GT, has_nt, nt_m, nt_v, first_head, GT, b, pair, split, tid,
                        NTHR, &s_ticket, nt_m + 4, nullptr, gs_tag);
}

// ------------------------------------------------------------------ host side
struct SplitPlan {
  int num_splits, tiles_per_split;
};

// OMX_DECODE_CLUSTER=0 turns the cluster / DSMEM combine off (A/B knob for the sweeps, not an API).
bool cluster_enabled() {
  static const bool on = [] {
    const char* e = getenv("OMX_DECODE_CLUSTER");
    return !e || atoi(e) != 0;
  }();
  return on;
}

// Can `cluster_x` CTAs of this kernel be co-scheduled as one cluster?  (> 8 needs the non-portable opt-in;
// a CTA that fills an SM's shared memory needs `cluster_x` free SMs in one GPC -- B200 GPCs have 16-20.)
template <typename K>
int cluster_capacity(K kern, dim3 grid, int threads, size_t smem, int cluster_x) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, size_t, int>, int> cache;
  int dev = 0;
  OMX_CUDA(cudaGetDevice(&dev));
  const auto key = std::make_tuple((const void*)kern, cluster_x, smem, dev);
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  bool ok = true;
  int n = 0;
  if (cluster_x > 8)
    ok = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
  if (ok) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster_x;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ok = cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess;
  }
  (void)cudaGetLastError();
  if (!ok) n = 0;
  if (getenv("OMX_DECODE_TRACE"))
    fprintf(stderr, "[omx decode] cluster size %d (%d threads, %zu B smem): %d co-resident clusters\n", cluster_x,
            threads, smem, n);
  cache[key] = n;
  return n;
}

// Which launches carry the programmatic-dependent-launch attribute: all of them.  CUDA-graph replays in one box
// (scripts/gpu_r02_pdl_ab.sh): the CUDA-core kernel gains from the overlapped dispatch alone (C1 15.56 -> 14.64 us);
// the TMA kernel -- one CTA fills an SM's shared memory -- only once its next launch fills its rings under the
// running launch's tail (stable_rows, see pdl_wait_prior_grid).  OMX_DECODE_PDL=0 / 1 forces the attribute off /
// on everywhere, OMX_DECODE_EARLY=0 keeps every read behind the wait (A/B knobs).
bool pdl_enabled(bool simt) {
  static const int forced = [] {
    const char* e = getenv("OMX_DECODE_PDL");
    return e ? (atoi(e) != 0 ? 1 : 0) : -1;
  }();
  (void)simt;
  return forced != 0;
}
bool early_reads_enabled() {
  static const bool on = [] {
    const char* e = getenv("OMX_DECODE_EARLY");
    return !e || atoi(e) != 0;
  }();
  return on;
}

template <typename K, typename... Args>
void launch_kernel(K kern, dim3 grid, int threads, size_t smem, cudaStream_t stream, int cluster_x, bool pdl,
                   Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl) {  // see pdl_wait_prior_grid()
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  OMX_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
}

// The cluster combine takes at most kMaxClusterSplits splits (OMX_DECODE_CLUSTER_MAX lowers the cap for
// sweeps): re-plan a wider split count down to it.
int cluster_cap() {
  static const int cap = [] {
    const char* e = getenv("OMX_DECODE_CLUSTER_MAX");
    const int v = e ? atoi(e) : kMaxClusterSplits;
    return std::max(2, std::min(v, kMaxClusterSplits));
  }();
  return cap;
}
SplitPlan clamp_for_cluster(SplitPlan sp, int n_tiles, int cap) {
  if (sp.num_splits <= cap) return sp;
  const int tps = (n_tiles + cap - 1) / cap;
  return {(n_tiles + tps - 1) / tps, tps};
}

// All-CTA combine (DecodeParams::gsync): one-wave grids whose splits are not combined through a cluster.
// OMX_DECODE_GSYNC=0 keeps the last-CTA combine (A/B knob).
constexpr int kMaxSplitsGsync = 160;  // > the SM count: a one-wave grid never has more CTAs per pair
bool gsync_enabled() {
  static const bool on = [] {
    const char* e = getenv("OMX_DECODE_GSYNC");
    return !e || atoi(e) != 0;
  }();
  return on;
}
// OMX_DECODE_GSLL=0: the all-CTA combine meets at a counter (gsync == 1) instead of polling tagged words (A/B knob)
bool gsll_enabled() {
  static const bool on = [] {
    const char* e = getenv("OMX_DECODE_GSLL");
    return !e || atoi(e) != 0;
  }();
  return on;
}
// does the slice bookkeeping of merge_and_store fit the merge scratch (n_ent x rows x D floats) and 4 items per warp?
bool gsync_fits(int splits, int n_heads, int D, int nthr, int scratch_floats) {
  const int D4 = D / 4, C = n_heads * D4;
  const int slice = (C + splits - 1) / splits;
  const GsyncShape gs = gsync_shape(splits);
  const int n_wi = gs.W == 32 ? slice * gs.K : (slice + 32 / gs.W - 1) / (32 / gs.W);
  const int nh = std::min(n_heads, (slice + D4 - 2) / D4 + 1);
  const int ml = (splits * nh + 3) & ~3;
  return n_wi <= 4 * (nthr / 32) && 2 * ml + 4 * slice * gs.K + slice <= scratch_floats;
}

// Split-K plan.  HBM bandwidth is a chip-wide resource, so what matters is (a) enough CTAs in flight
// to cover it -- about one per SM, each with >= 96 KB of TMA loads outstanding -- and (b) as little
// per-CTA overhead (q staging, partial write, combine) as possible.  Measured on B200
// (gpurun_out/s19_sweep.log): C2 (512 (batch, kv-head) pairs) is fastest with NO split (299 us vs 315 us
// with 4); C5 (8 pairs) is fastest with one CTA per SM (18 splits: 31 us; 37 splits: 40 us; 9: 33 us).
SplitPlan plan_splits(int64_t pairs, int n_tiles, int sms, int min_tiles, int cap = kMaxSplits) {
  if (n_tiles <= 0) return {1, 1};
  static const int forced = [] {  // debugging / tuning knob, not an API
    const char* e = getenv("OMX_DECODE_SPLITS");
    return e ? atoi(e) : 0;
  }();
  int want = forced > 0 ? forced : (pairs >= sms ? 1 : (int)(sms / pairs));
  const int max_s = std::max(1, std::min(n_tiles / std::max(1, min_tiles), cap));
  want = std::max(1, std::min(want, forced > 0 ? std::min(n_tiles, cap) : max_s));
  const int tps = (n_tiles + want - 1) / want;
  return {(n_tiles + tps - 1) / tps, tps};
}

bool inner_contig(const omx_array* a) { return a->strides[3] == 1 || a->shape[3] == 1; }

template <typename T, int VE>
void launch_simt(DecodeParams& p, cudaStream_t stream, int Gt, dim3 grid, bool one_wave, bool want_cluster) {
  // staged variant: the split's rows as two contiguous blocks behind the kernel's own shared memory
  const size_t stage_bytes = 2 * (size_t)p.tiles_per_split * kTile * 32 * VE * sizeof(T);
  auto base_smem = [&](int gt) {
    return sizeof(float) * ((size_t)kSimtWarps * gt * (32 * VE + 2) + 2 * 32 * VE + 4 + cluster_scratch_floats(gt)) +
           sizeof(T) * (size_t)gt * 32 * VE;
  };
  static const int staged_env = [] {  // OMX_DECODE_STAGED=0: never (A/B knob)
    const char* e = getenv("OMX_DECODE_STAGED");
    return e ? atoi(e) : 1;
  }();
  const bool rows_contig = p.ks[2] == 32 * VE && p.vs[2] == 32 * VE && p.ks[3] == 1 && p.vs[3] == 1;
  constexpr bool kHasStaged = std::is_same<T, float>::value && VE == 4;  // instantiated where the 16-key body is
  const bool staged = kHasStaged && staged_env != 0 && one_wave && !p.paged && rows_contig &&
                      ((base_smem(4) + 127) & ~(size_t)127) + stage_bytes <= 200 * 1024;
  auto go = [&](auto kern, int gt) {
    size_t smem = base_smem(gt);
    p.stage_off = p.stage_keys = 0;
    if (staged) {
      p.stage_off = (int)((smem + 127) & ~(size_t)127);
      p.stage_keys = p.tiles_per_split * kTile;
      smem = (size_t)p.stage_off + stage_bytes;
    }
    OMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.cluster = want_cluster && p.num_splits > 1 &&
                        cluster_capacity(kern, grid, kSimtWarps * 32, smem, p.num_splits) >= (int)(grid.y * grid.z)
                    ? 1 : 0;
    const int64_t ctas = (int64_t)grid.x * grid.y * grid.z;
    p.gsync = (!p.cluster && p.num_splits > 1 && p.counters2 && gsync_enabled() && ctas <= sm_count() &&
               gsync_fits(p.num_splits, gt, 32 * VE, kSimtWarps * 32, kSimtWarps * gt * 32 * VE))
                  ? 1 : 0;
    if (p.gsync && p.ws_w) p.gsync = 2;
    if (p.gsync) p.peer_total = (int)ctas;
    launch_kernel(kern, grid, kSimtWarps * 32, smem, stream, p.cluster ? p.num_splits : 1, pdl_enabled(true), p);
  };
  // the 16-key variant keeps 2 x 16 rows per lane in registers; instantiated where it is used and measured --
  // float32 at head_dim 128 (C1; 16-bit head_dim 128 runs on the TMA kernel) -- to keep the build time down
  constexpr int KBIG = kHasStaged ? 16 : 4;
  if (staged && KBIG != 4) {
    switch (Gt) {
      case 1: go(decode_simt_kernel<T, VE, 1, 4, true>, 1); break;
      case 2: go(decode_simt_kernel<T, VE, 2, 4, true>, 2); break;
      default: go(decode_simt_kernel<T, VE, 4, 4, true>, 4); break;
    }
  } else if (one_wave && KBIG != 4) {
    switch (Gt) {
      case 1: go(decode_simt_kernel<T, VE, 1, KBIG>, 1); break;
      case 2: go(decode_simt_kernel<T, VE, 2, KBIG>, 2); break;
      default: go(decode_simt_kernel<T, VE, 4, KBIG>, 4); break;
    }
  } else {
    switch (Gt) {
      case 1: go(decode_simt_kernel<T, VE, 1, 4>, 1); break;
      case 2: go(decode_simt_kernel<T, VE, 2, 4>, 2); break;
      default: go(decode_simt_kernel<T, VE, 4, 4>, 4); break;
    }
  }
  count_launch();
  OMX_CUDA(cudaGetLastError());
}

template <typename T>
void launch_simt_d(DecodeParams& p, cudaStream_t stream, int Gt, dim3 grid, bool one_wave, bool want_cluster) {
  switch (p.D) {
    case 32: launch_simt<T, 1>(p, stream, Gt, grid, one_wave, want_cluster); break;
    case 64: launch_simt<T, 2>(p, stream, Gt, grid, one_wave, want_cluster); break;
    case 128: launch_simt<T, 4>(p, stream, Gt, grid, one_wave, want_cluster); break;
    default: launch_simt<T, 8>(p, stream, Gt, grid, one_wave, want_cluster); break;
  }
}

__global__ void peer_wait_kernel(const unsigned* flags, int world, unsigned expected, int rank) {
  if ((int)threadIdx.x >= world) return;
  // expected == 0: "as many arrivals as this rank has itself signalled" -- its own decode launch (earlier on the
  // stream) bumped flags[rank], so the count needs no host argument and the wait can be replayed from a graph
  if (expected == 0) expected = *(const volatile unsigned*)(flags + rank);
  const volatile unsigned* f = flags + threadIdx.x;
  unsigned spins = 0;
  // counters only grow; the signed difference tolerates wrap-around
  while ((int)(*f - expected) < 0) {
    __nanosleep(64);
    if (++spins > (1u << 25)) __trap();  // a lost peer becomes a launch failure, not a hung GPU
  }
  __threadfence_system();
}

// The data + flag exchange as a kernel of its own (launch shapes whose final store is not the all-CTA combine):
// one CTA sends this rank's [B,Hq,D] slice of out_full as staging words and unpacks every peer's slice.
template <typename T>
__global__ void ll_exchange_kernel(dd::LLDev ll, T* out_full, int64_t os0, int64_t os1, int B, int Hq, int D) {
  using LP = dd::LLPack<T>;
  const unsigned seq = *ll.seq + 1u;
  const int quads = B * Hq * D / 4;
  for (int qi = threadIdx.x; qi < quads; qi += blockDim.x) {
    const int64_t e = (int64_t)qi * 4;
    const int b = (int)(e / ((int64_t)Hq * D)), h = (int)((e / D) % Hq), d = (int)(e % D);
    uint32_t w[LP::NW];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(out_full + b * os0 + ((int64_t)ll.rank * Hq + h) * os1 + d);
#pragma unroll
    for (int j = 0; j < LP::NW; ++j) w[j] = src[j];
    dd::ll_exchange<LP::NW>(ll, seq, e * (int64_t)sizeof(T) / 4, w, [&](int r, const uint32_t (&x)[LP::NW]) {
      LP::store(out_full + b * os0 + ((int64_t)r * Hq + h) * os1 + d, x);
    });
  }
  __syncthreads();
  if (threadIdx.x == 0) *ll.seq = seq;
}

// One CTA per (batch, head): wait for every rank's arrival, then the log-sum-exp merge of the partial slots.
template <typename T>
__global__ void seqshard_merge_kernel(T* out, int64_t os0, int64_t os1, int64_t os3, const float* partial, int world,
                                      int Hq, int D, const unsigned* flags, unsigned expected, int rank) {
  __shared__ float s_w[kMaxPeers];
  const int bh = blockIdx.x, b = bh / Hq, h = bh % Hq;
  if (flags) {
    if ((int)threadIdx.x < world) {
      if (expected == 0) expected = *(const volatile unsigned*)(flags + rank);  // see peer_wait_kernel
      const volatile unsigned* f = flags + threadIdx.x;
      unsigned spins = 0;
      while ((int)(*f - expected) < 0) {  // counters only grow; the signed difference tolerates wrap-around
        __nanosleep(64);
        if (++spins > (1u << 25)) __trap();  // a lost peer becomes a launch failure, not a hung GPU
      }
      __threadfence_system();
    }
    __syncthreads();
  }
  const int64_t slot = (int64_t)D + 2;
  const float* base = partial + ((int64_t)b * Hq + h) * slot;
  const int64_t rank_stride = (int64_t)gridDim.x * slot;
  if (threadIdx.x == 0) {
    float M = -INFINITY;
    for (int r = 0; r < world; ++r) M = fmaxf(M, __ldcg(base + r * rank_stride + D));
    float W = 0.f;
    for (int r = 0; r < world; ++r) {
      const float m = __ldcg(base + r * rank_stride + D), l = __ldcg(base + r * rank_stride + D + 1);
      const float w = (m > -INFINITY && l > 0.f) ? l * fast_exp2(m - M) : 0.f;
      s_w[r] = w;
      W += w;
    }
    const float inv = 1.0f / W;
    for (int r = 0; r < world; ++r) s_w[r] *= inv;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < world; ++r) {
      const float w = s_w[r];
      if (w > 0.f) acc = fmaf(__ldcg(base + r * rank_stride + d), w, acc);  // a rank without keys holds no number
    }
    out[b * os0 + h * os1 + d * os3] = Num<T>::from_f(acc);
  }
}

}  // namespace

#if OMX_DECODE_PART == 1
void dd::launch_simt_bf16(DecodeParams& p, cudaStream_t stream, int Gt, dim3 grid, bool one_wave, bool want_cluster) {
  launch_simt_d<__nv_bfloat16>(p, stream, Gt, grid, one_wave, want_cluster);
}
#elif OMX_DECODE_PART == 2
void dd::launch_simt_f16(DecodeParams& p, cudaStream_t stream, int Gt, dim3 grid, bool one_wave, bool want_cluster) {
  launch_simt_d<__half>(p, stream, Gt, grid, one_wave, want_cluster);
}
#elif OMX_DECODE_PART == 3
void dd::launch_simt_f32(DecodeParams& p, cudaStream_t stream, int Gt, dim3 grid, bool one_wave, bool want_cluster) {
  launch_simt_d<float>(p, stream, Gt, grid, one_wave, want_cluster);
}
#endif

#if OMX_DECODE_PART == 0

void seqshard_merge(const omx_array* out, const float* partial, int world, int B, int Hq, int D,
                    const unsigned* flags, unsigned expected, int rank, cudaStream_t stream) {
  if (B * Hq == 0) return;
  const int threads = std::min(128, std::max(32, D));
  auto go = [&](auto* o) {
    using T = std::remove_pointer_t<decltype(o)>;
    seqshard_merge_kernel<T><<<B * Hq, threads, 0, stream>>>(o, out->strides[0], out->strides[1], out->strides[3],
                                                            partial, world, Hq, D, flags, expected, rank);
  };
  note_launch("seqshard_merge");
  switch (out->dtype) {
    case OMX_FLOAT32: go((float*)out->data); break;
    case OMX_BFLOAT16: go((__nv_bfloat16*)out->data); break;
    default: go((__half*)out->data); break;
  }
  count_launch();
  OMX_CUDA(cudaGetLastError());
}

void peer_wait(const unsigned* flags, int world, unsigned expected, int rank, cudaStream_t stream) {
  peer_wait_kernel<<<1, 32, 0, stream>>>(flags, world, expected, rank);
  count_launch();
  OMX_CUDA(cudaGetLastError());
}

size_t decode_graph_scratch_bytes(int B, int Hkv, int Hq, int D, int dtype, int max_rows) {
  // mirrors the two plans of decode_attention at the largest position the launch may see
  const int sms = sm_count();
  const int n_tiles = (std::max(max_rows - 1, 0) + kTile - 1) / kTile;
  const int G = Hq / std::max(Hkv, 1);
  size_t worst = 0;
  for (int simt = 0; simt < 2; ++simt) {
    if (!simt && !(dtype != OMX_FLOAT32 && D == 128 && G <= 16)) continue;
    const int Gt = simt ? ((G % 4 == 0) ? 4 : (G % 2 == 0 ? 2 : 1)) : G;
    const int64_t pairs = (int64_t)B * Hkv * (simt ? G / Gt : 1);
    const SplitPlan sp = plan_splits(pairs, n_tiles, sms, simt ? 2 : 4,
                                     (!simt && gsync_enabled()) ? kMaxSplitsGsync : kMaxSplits);
    const size_t part = sp.num_splits > 1 ? (size_t)pairs * sp.num_splits * Gt * (D + 2) : 0;
    worst = std::max(worst, (simt ? sizeof(float) : 8) * part + sizeof(int) * (2 * (size_t)pairs + 2));  // (8: tagged words)
  }
  return worst;
}

bool decode_supported(const SdpaArgs& a, const char** why) {
  auto no = [&](const char* w) {
    if (why) *why = w;
    return false;
  };
  if (a.Lq != 1) return no("Lq != 1");
  if (a.mask_mode == MASK_ADD && a.mask->dtype != a.q->dtype) return no("additive mask dtype differs from q");
  if (a.D != a.Dv) return no("Dk != Dv");
  if (!(a.D == 32 || a.D == 64 || a.D == 128 || a.D == 256)) return no("head_dim not in {32,64,128,256}");
  if (!inner_contig(a.q) || !inner_contig(a.k) || !inner_contig(a.v) || !inner_contig(a.out))
    return no("innermost axis not contiguous");
  const size_t es = dtype_size(a.q->dtype);
  const int64_t v = (int64_t)(16 / es);
  if (!aligned16(a.k->data) || !aligned16(a.v->data)) return no("K/V base not 16-byte aligned");
  for (int i = 0; i < 3; ++i)
    if (a.k->strides[i] % v || a.v->strides[i] % v) return no("K/V strides not multiples of 16 bytes");
  if (a.out->dtype != a.q->dtype) return no("out dtype differs");
  return true;
}

namespace {
// OMX_DECODE_TRACE=1: every decode launch records per-CTA phase timestamps, is synchronised and dumped to
// stderr (debugging aid for the latency-bound single-sequence shapes; never set in production).
struct TraceDump {
  unsigned long long* dev = nullptr;
  size_t ctas = 0;
  cudaStream_t stream = nullptr;
  ~TraceDump() {
    if (!dev) return;
    cudaStreamSynchronize(stream);
    constexpr int NS = 16;
    std::vector<unsigned long long> h(ctas * NS);
    cudaMemcpy(h.data(), dev, h.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(dev);
    unsigned long long t0 = ~0ull, t1 = 0;
    for (size_t c = 0; c < ctas; ++c)
      if (h[c * NS]) t0 = std::min(t0, h[c * NS]);
    double sum[NS] = {0}, mx[NS] = {0};
    int cnt[NS] = {0};
    for (size_t c = 0; c < ctas; ++c)
      for (int s = 0; s < NS; ++s)
        if (h[c * NS + s]) {
          const double us = (double)(h[c * NS + s] - t0) * 1e-3;
          sum[s] += us; mx[s] = std::max(mx[s], us); ++cnt[s];
          t1 = std::max(t1, h[c * NS + s]);
        }
    fprintf(stderr, "[omx decode trace] ctas=%zu span=%.2fus | slot: n mean max (us since first CTA start):", ctas,
            (double)(t1 - t0) * 1e-3);
    static const char* nm[NS] = {"start", "q_staged", "loop_done", "merged", "fenced", "ticket", "end", "tma_issued",
                                 "c_fence", "c_ml", "c_weights", "c_folded", "p_loaded", "p_rms", "p_issued", ""};
    static const int order[NS] = {0, 7, 14, 12, 13, 1, 2, 3, 4, 5, 8, 9, 10, 11, 6, 15};
    for (int i = 0; i < NS; ++i) {
      const int s = order[i];
      if (cnt[s]) fprintf(stderr, " %s: %d %.2f %.2f |", nm[s], cnt[s], sum[s] / cnt[s], mx[s]);
    }
    fprintf(stderr, "\n");
    const char* e = getenv("OMX_DECODE_TRACE");
    if (e && atoi(e) > 1)  // every CTA: index, then the slots in us since the first start (0 = not reached)
      for (size_t c = 0; c < ctas; ++c) {
        fprintf(stderr, "[omx decode trace cta %zu]", c);
        for (int s = 0; s < NS; ++s)
          fprintf(stderr, " %.2f", h[c * NS + s] ? (double)(h[c * NS + s] - t0) * 1e-3 : 0.0);
        fprintf(stderr, "\n");
      }
  }
};
}  // namespace

void decode_attention(const SdpaArgs& a, const DecodeFused& f, cudaStream_t stream) {
  DecodeParams p{};
  static const bool trace_on = [] {
    const char* e = getenv("OMX_DECODE_TRACE");
    return e && atoi(e) > 0;
  }();
  TraceDump trace;  // destroyed after the launch below
  auto arm_trace = [&](dim3 grid) {
    if (!trace_on) return;
    trace.ctas = (size_t)grid.x * grid.y * grid.z;
    trace.stream = stream;
    OMX_CUDA(cudaMalloc(&trace.dev, trace.ctas * 128));
    OMX_CUDA(cudaMemset(trace.dev, 0, trace.ctas * 128));
    p.trace = trace.dev;
  };
  p.q = a.q->data;
  p.out = a.out->data;
  p.k = a.k->data;
  p.v = a.v->data;
  for (int i = 0; i < 4; ++i) {
    p.qs[i] = a.q->strides[i];
    p.os[i] = a.out->strides[i];
    p.ks[i] = a.k->strides[i];
    p.vs[i] = a.v->strides[i];
  }
  p.B = a.B; p.Hq = a.Hq; p.Hkv = a.Hkv; p.G = a.Hq / a.Hkv; p.D = a.D;
  p.Lk = a.Lk;
  p.fused = f.enabled ? 1 : 0;
  p.append = f.enabled && f.append ? 1 : 0;
  p.partial = f.partial ? 1 : 0;
  p.n_mem = p.append ? a.Lk - 1 : a.Lk;
  p.scale_log2 = a.scale * kLog2e;
  if (f.enabled) {
    if (f.append) {
      p.k_new = f.k_new->data;
      p.v_new = f.v_new->data;
    }
    for (int i = 0; i < 4; ++i) {
      p.kns[i] = f.append ? f.k_new->strides[i] : 0;
      p.vns[i] = f.append ? f.v_new->strides[i] : 0;
      p.kcs[i] = a.k->strides[i];
      p.vcs[i] = a.v->strides[i];
    }
    p.k_row0 = a.k->data;
    p.v_row0 = a.v->data;
    p.rope_dims = f.rope_dims;
    p.traditional = f.traditional ? 1 : 0;
    if (f.rope_dims > 0) {
      p.cos_row = f.table.cos + (size_t)f.position * f.table.half;
      p.sin_row = f.table.sin + (size_t)f.position * f.table.half;
    }
    p.q_norm_w = f.q_norm_w;
    p.k_norm_w = f.k_norm_w;
    p.norm_eps = f.norm_eps;
    p.norm_inv_n = 1.0f / (float)a.D;
  }
  if (a.B == 0 || a.Hq == 0) return;
  OMX_CHECK(p.Lk >= 1, "[scaled_dot_product_attention] decode needs at least one key");
  if (f.peers) {
    p.n_peers = f.peers->world;
    p.peer_rank = f.peers->rank;
    p.peer_wait = f.peer_wait ? 1 : 0;
    for (int r = 0; r < p.n_peers; ++r) {  // already shifted to this rank's first head by the caller
      p.peer_out[r] = f.peers->out[r];
      p.peer_flag[r] = f.peers->flags[r];
    }
  }

  if (f.ll) {  // data + flag exchange: armed below if the launch ends in the all-CTA combine, else a second kernel
    OMX_CHECK(!f.peers, "[attn_decode_fused_sharded_ll] one exchange protocol per launch");
    OMX_CHECK(a.out->strides[3] == 1 && a.out->strides[0] % 4 == 0 && a.out->strides[1] % 4 == 0 &&
                  ((uintptr_t)f.ll_out_full & 15) == 0 && ((uintptr_t)a.out->data & 15) == 0 && a.D % 4 == 0,
              "[attn_decode_fused_sharded_ll] out_full rows must be contiguous and 16-byte aligned");
  }
  auto ll_dev = [&]() {
    dd::LLDev d{};
    d.world = f.ll->world;
    d.rank = f.ll->rank;
    for (int r = 0; r < d.world; ++r) d.buf[r] = (unsigned long long*)f.ll->staging[r];
    d.seq = f.ll->seq;
    d.words = (int)((int64_t)a.B * a.Hq * a.D * (int64_t)dtype_size(a.q->dtype) / 4);
    return d;
  };
  auto ll_second_kernel = [&]() {  // the launch stored the local slice plainly: exchange it now
    const dd::LLDev d = ll_dev();
    auto go = [&](auto* o) {
      using T = std::remove_pointer_t<decltype(o)>;
      ll_exchange_kernel<T><<<1, 256, 0, stream>>>(d, o, a.out->strides[0], a.out->strides[1], a.B, a.Hq, a.D);
    };
    switch (a.q->dtype) {
      case OMX_FLOAT32: go((float*)f.ll_out_full); break;
      case OMX_BFLOAT16: go((__nv_bfloat16*)f.ll_out_full); break;
      default: go((__half*)f.ll_out_full); break;
    }
    count_launch();
    OMX_CUDA(cudaGetLastError());
  };

  const bool masked = a.mask_mode == MASK_BOOL || a.mask_mode == MASK_ADD;
  if (masked) {
    OMX_CHECK(!f.enabled, "the fused decode step takes no array mask");
    p.mask = a.mask->data;
    p.mask_kind = a.mask_mode == MASK_BOOL ? 1 : 2;
    p.mks[0] = a.mask_strides[0];
    p.mks[1] = a.mask_strides[1];
    p.mks[2] = a.mask_strides[3];
  }
  const bool paged = f.enabled && f.paged != nullptr;
  const bool dyn = f.enabled && f.pos_dev != nullptr && !paged;  // graph mode: cache-owned scratch, fixed addresses
  if (dyn) {
    OMX_CHECK(!f.peers && !masked, "the dynamic-position decode step takes no peer group and no array mask");
    OMX_CHECK(f.max_rows >= 1 && a.Lk == f.max_rows, "dynamic-position decode: K/V views must span max_rows");
    p.pos_dev = f.pos_dev;
    p.pos_stride = 0;
    p.max_rows = f.max_rows;
  }
  if (f.enabled && !dyn && !paged && early_reads_enabled() && pdl_enabled(false))
    p.stable_rows = std::max(0, std::min(f.stable_rows, p.n_mem));
  {
    // sweep knobs (scripts/gpu_r02_l2_ab*.sh).  In-loop L2 requests measured 10 % SLOWER at every batch (a tile
    // asked into L2 and then loaded evict-first costs more than it hides), so l2_ahead stays 0; the pre-wait
    // depth is set per grid type below (l2_early < 0 here = default).
    static const int ahead = [] { const char* e = getenv("OMX_DECODE_L2AHEAD"); return e ? atoi(e) : 0; }();
    static const int early = [] { const char* e = getenv("OMX_DECODE_L2EARLY"); return e ? atoi(e) : -1; }();
    p.l2_ahead = std::max(0, ahead);
    p.l2_early = early;
  }
  if (paged) {
    // a.k / a.v describe the POOL: data = base, strides[0] = page stride, [1] = head stride inside a page,
    // [2] = row stride; a.Lk = longest sequence after the step (host mirror) -- it sizes the split plan only,
    // every CTA derives its own key count from lens[b]
    OMX_CHECK(!f.peers && !masked, "the paged decode step takes no peer group and no array mask");
    p.paged = 1;
    p.block_table = f.paged->block_table;
    p.bt_stride = f.paged->bt_stride;
    p.pos_dev = f.paged->lens_in;
    p.pos_stride = 1;
    p.lens_out = f.paged->lens_out;
    p.max_rows = f.paged->bt_stride * kTile;
  }
  // one workspace request per call (a second one could move the first): split-K partials, then flags.
  // Graph mode carves the same layout (+ the counters) out of the cache-owned scratch instead, whose
  // address never changes under a captured launch.
  // n_words > 0: the tagged-word variant of the all-CTA combine (gsync == 2) -- its partials live in a pool of
  // their own (only ever {value, tag} words, zeroed when allocated, tags only grow) next to the launch counter
  auto carve_workspace = [&](size_t no, size_t nml, size_t n_ctr, size_t n_words = 0) {
    if (dyn || (paged && f.scratch)) {  // cache-owned scratch: its address is stable under a captured launch
      const size_t need = n_words ? 8 * n_words + sizeof(int) * (n_ctr + 1)
                                  : sizeof(float) * (no + nml) + sizeof(int) * n_ctr;
      OMX_CHECK(need <= f.scratch_bytes, "dynamic-position decode: scratch too small (%zu > %zu bytes); call "
                "omx_kv_cache_prepare_graph with this launch's head count first", need, f.scratch_bytes);
      float* ws = (float*)f.scratch;
      if (n_words) {  // (graph mode only: one plan per cache, so the layout -- and the counter's place -- never moves)
        p.ws_w = reinterpret_cast<uint2*>(ws);
        p.counters = reinterpret_cast<int*>(p.ws_w + n_words);
        p.peer_done = p.counters + n_ctr - 1;
        p.gs_seq = reinterpret_cast<unsigned*>(p.counters + n_ctr);
        return;
      }
      p.ws_o = ws;
      p.ws_ml = ws + no;
      if (n_ctr) {
        p.counters = reinterpret_cast<int*>(ws + no + nml);
        p.peer_done = p.counters + n_ctr - 1;
      }
      return;
    }
    const size_t fl = masked ? (size_t)a.B * a.Hq : 0;
    if (n_words) {
      p.ws_w = reinterpret_cast<uint2*>(get_tagged_workspace(8 * n_words, stream, &p.gs_seq));
      no = nml = 0;
    }
    if (no + nml + fl) {
      float* ws = (float*)get_workspace(sizeof(float) * (no + nml) + fl, stream);
      if (no) {
        p.ws_o = ws;
        p.ws_ml = ws + no;
      }
      if (fl) p.dead = reinterpret_cast<uint8_t*>(ws + no + nml);
    }
    if (n_ctr) {
      p.counters = get_counters(n_ctr, stream);
      p.peer_done = p.counters + n_ctr - 1;
    }
  };

  const int sms = sm_count();
  const int n_tiles = (p.n_mem + kTile - 1) / kTile;
  const bool b16 = a.q->dtype != OMX_FLOAT32;
  const bool use_hmma = b16 && a.D == 128 && p.G <= 16 && a.k->strides[2] >= 128 && a.v->strides[2] >= 128;

  if (use_hmma) {
    // cfg 0: 3 stages x 2 CTAs/SM (default); cfg 1: 6 stages x 1 CTA/SM.  OMX_DECODE_CFG is a
    // tuning knob for the bench sweeps, not an API.
    static const int cfg_env = [] {
      const char* e = getenv("OMX_DECODE_CFG");
      return e ? atoi(e) : -1;
    }();
    const int64_t pairs = (int64_t)a.B * a.Hkv;
    // all-CTA combine: a one-wave grid whose splits no cluster takes; it also lifts the 64-split cap (ONE pair --
    // a rank of the kv-head-sharded C5 -- spreads over 128 CTAs of 4 tiles instead of 64 of 8)
    auto gsync_for = [&](const SplitPlan& c) {
      return gsync_enabled() && c.num_splits > 1 && pairs * c.num_splits <= sms &&
             gsync_fits(c.num_splits, p.G, 128, 7 * 32, 6 * 16 * 128);
    };
    SplitPlan natural = plan_splits(pairs, n_tiles, sms, 4, gsync_enabled() ? kMaxSplitsGsync : kMaxSplits);
    if (natural.num_splits > kMaxSplits && !gsync_for(natural)) natural = plan_splits(pairs, n_tiles, sms, 4);
    const bool want_cluster = cluster_enabled() && !masked && !f.peers && !f.ll && natural.num_splits > 1;
    // Cluster policy (measured, scripts/gpu_cluster_sweep.sh): the DSMEM combine saves ~2 us of tail, but a
    // cluster must be co-resident in one GPC -- with one CTA per SM this B200 places 8 clusters of <= 10
    // CTAs (7 of 16).  So: the natural plan if it fits; else a plan clamped to kClusterClamp splits when
    // that costs at most 6 more 64-key tiles per CTA (Qwen3-8B ctx 8192: 17.6 -> 16.2 us; at ctx 32768
    // the lost SMs cost more than the tail); else the HBM/L2 combine.
    constexpr int kClusterClamp = 10;
    static const bool cap_from_env = getenv("OMX_DECODE_CLUSTER_MAX") != nullptr;  // sweep knob, read once
    const int cap = cap_from_env ? cluster_cap() : kClusterClamp;
    const SplitPlan clamped = clamp_for_cluster(natural, n_tiles, cap);
    const bool try_clamped = clamped.num_splits != natural.num_splits &&
                             clamped.tiles_per_split - natural.tiles_per_split <= 6;
    // a grid that fits one CTA per SM runs the 6-stage / 1-CTA-per-SM variant (deeper TMA pipeline per
    // CTA: C5 31 us vs 34 us); larger grids the 3-stage / 2-CTA-per-SM one (C2 299 us vs 307 us)
    const bool deep = cfg_env == 1 || (cfg_env < 0 && pairs * natural.num_splits <= sms);
    const int cfg = deep ? 1 : 0;
    const int NSTAGE = cfg == 1 ? 6 : 3;  // consumer warps == stages (see kernel comment)
    // tiles past the ring asked into L2 before the dependency wait: ~5 us of the previous launch's tail x HBM
    // rate = 8 tiles per CTA on one-wave grids (8 rows x 8 kv heads: 44.6 -> 42.9 us, 16 rows: 80.6 -> 79.0);
    // multi-wave grids only overlap their last wave (32 rows: 160.2 -> 157.9, 64 rows: 305.7 -> 303.7)
    if (p.l2_early < 0) p.l2_early = deep ? 8 : 4;
    p.peer_total = (int)pairs;
    const bool bf = a.q->dtype == OMX_BFLOAT16;
    // graph mode: the map spans every reserved row (the tail beyond the position is masked in the kernel)
    // paged: the map spans the pool, [n_pages][Hkv][64][128]; one box = half the features of a whole page
    const uint64_t rows = paged ? (uint64_t)kTile : (uint64_t)std::max(dyn ? p.max_rows : p.n_mem, 1);
    const uint64_t outer = paged ? (uint64_t)f.paged->n_pages : (uint64_t)a.B;
    CUtensorMap tmK = make_tmap_4d_b16(a.k->data, 128, rows, a.Hkv, outer, a.k->strides[2], a.k->strides[1],
                                       a.k->strides[0], 64, 64, bf);
    CUtensorMap tmV = make_tmap_4d_b16(a.v->data, 128, rows, a.Hkv, outer, a.v->strides[2], a.v->strides[1],
                                       a.v->strides[0], 64, 64, bf);
    const size_t smem_base = 1024 + (size_t)NSTAGE * kStageBytes + 16 * kQPitch * 2 + sizeof(float) * (128 + 128 + 16);
    // few-way cluster splits fold through receive slots in rank 0 (push combine): a little more shared memory
    static const bool push_env = [] {  // OMX_DECODE_PUSH=0: always the pull combine (A/B knob)
      const char* e = getenv("OMX_DECODE_PUSH");
      return !e || atoi(e) != 0;
    }();
    const bool push_ok = push_env && want_cluster && natural.num_splits <= kPushSplits && p.G <= 8 && !f.partial;
    const size_t smem = smem_base + (push_ok ? sizeof(float) * push_recv_floats(natural.num_splits, p.G, 128) : 0);
    auto go = [&](auto kern) {
      OMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int threads = (NSTAGE + 1) * 32;
      auto fits = [&](const SplitPlan& c) {
        return c.num_splits > 1 && c.num_splits <= kMaxClusterSplits &&
               cluster_capacity(kern, dim3(c.num_splits, a.Hkv, a.B), threads, smem, c.num_splits) >= pairs;
      };
      SplitPlan sp = natural;
      p.cluster = 0;
      if (want_cluster) {
        if (fits(natural)) p.cluster = 1;
        else if (try_clamped && fits(clamped)) { sp = clamped; p.cluster = 1; }
      }
      static const bool gsync_first = [] {  // OMX_DECODE_GSYNC=2: all-CTA combine even where a clamped cluster fits
        const char* e = getenv("OMX_DECODE_GSYNC");
        return e && atoi(e) == 2;
      }();
      if (gsync_first && p.cluster && sp.num_splits != natural.num_splits && gsync_for(natural)) {
        sp = natural;
        p.cluster = 0;
      }
      p.num_splits = sp.num_splits;
      p.tiles_per_split = sp.tiles_per_split;
      p.push_combine = (push_ok && p.cluster && p.num_splits <= kPushSplits) ? 1 : 0;
      p.gsync = (!p.cluster && deep && gsync_for(sp)) ? 1 : 0;
      // tagged words instead of the meeting point: not with the r01 peer counters (they share the launch-wide
      // counter) and not on the paged cache's scratch (its plan, hence its layout, changes as the sequences grow)
      // ... and only for short key loops (<= 16 tiles per CTA): there the CTAs of a pair finish together and the
      // polls hit at once (one rank of the sharded C5 14.3 -> 13.2 us, C1 13.9 -> 13.0); behind a long loop the
      // early CTAs' polling competes with the late CTAs' streaming (C5 on one GPU, 29 tiles: 29.0 -> 29.5 us)
      if (p.gsync && gsll_enabled() && !p.n_peers && !paged && sp.tiles_per_split <= 16) p.gsync = 2;
      if (p.gsync) p.peer_total = (int)pairs * p.num_splits;  // every CTA stores a slice
      if (p.gsync && f.ll) {
        p.ll = ll_dev();
        p.ll_out = f.ll_out_full;
      }
      carve_workspace(p.num_splits > 1 ? (size_t)pairs * p.num_splits * p.G * p.D : 0,
                      p.num_splits > 1 ? (size_t)pairs * p.num_splits * p.G * 2 : 0,
                      (p.num_splits > 1 || p.n_peers) ? (size_t)(p.gsync ? 2 : 1) * pairs + 1 : 0,
                      p.gsync == 2 ? (size_t)pairs * p.num_splits * p.G * (p.D + 2) : 0);
      p.counters2 = p.gsync ? p.counters + pairs : nullptr;
      dim3 grid(p.num_splits, a.Hkv, a.B);
      arm_trace(grid);
      launch_kernel(kern, grid, threads, smem, stream, p.cluster ? p.num_splits : 1, pdl_enabled(false), tmK, tmV, p);
    };
    note_launch("decode_hmma_tma");
    if (bf) {
      if (cfg == 1) {
        if (p.G > 8) go(decode_hmma_kernel<__nv_bfloat16, 6, true, 1>);
        else go(decode_hmma_kernel<__nv_bfloat16, 6, false, 1>);
      } else {
        if (p.G > 8) go(decode_hmma_kernel<__nv_bfloat16, 3, true, 2>);
        else go(decode_hmma_kernel<__nv_bfloat16, 3, false, 2>);
      }
    } else {
      if (cfg == 1) {
        if (p.G > 8) go(decode_hmma_kernel<__half, 6, true, 1>);
        else go(decode_hmma_kernel<__half, 6, false, 1>);
      } else {
        if (p.G > 8) go(decode_hmma_kernel<__half, 3, true, 2>);
        else go(decode_hmma_kernel<__half, 3, false, 2>);
      }
    }
    count_launch();
    OMX_CUDA(cudaGetLastError());
    if (masked) masked_rows_fixup(a, p.dead, stream);
    if (f.ll && !p.ll.world) ll_second_kernel();
    return;
  }

  // ---- CUDA-core path
  const int Gt = (p.G % 4 == 0) ? 4 : (p.G % 2 == 0 ? 2 : 1);
  const int groups = p.G / Gt;
  const int64_t pairs = (int64_t)a.B * a.Hkv * groups;
  SplitPlan sp = plan_splits(pairs, n_tiles, sms, 2);
  // CUDA-core kernel: cluster combine only when the natural plan is co-resident as it is (a clamped plan
  // lost more in the key loop than the combine saved: C1 12.9 -> 13.6 us)
  const bool want_cluster = cluster_enabled() && !masked && !f.peers && !f.ll && sp.num_splits > 1 &&
                            sp.num_splits <= kMaxClusterSplits;
  p.num_splits = sp.num_splits;
  p.tiles_per_split = sp.tiles_per_split;
  // (room for the all-CTA combine's second counter row; launch_simt decides once it knows about the cluster)
  carve_workspace(p.num_splits > 1 ? (size_t)pairs * p.num_splits * Gt * p.D : 0,
                  p.num_splits > 1 ? (size_t)pairs * p.num_splits * Gt * 2 : 0,
                  (p.num_splits > 1 || p.n_peers) ? (size_t)2 * pairs + 1 : 0);
  p.counters2 = p.counters ? p.counters + pairs : nullptr;
  // the tagged-word pool for the all-CTA combine, should launch_simt pick it (plain launches only: graph / paged
  // launches of this kernel keep the meeting point)
  if (p.num_splits > 1 && gsll_enabled() && !p.n_peers && !dyn && !paged && pairs * p.num_splits <= sms)
    p.ws_w = reinterpret_cast<uint2*>(
        get_tagged_workspace(8 * (size_t)pairs * p.num_splits * Gt * (p.D + 2), stream, &p.gs_seq));
  p.peer_total = (int)pairs;
  dim3 grid(p.num_splits, a.Hkv * groups, a.B);
  arm_trace(grid);
  note_launch("decode_simt");
  static const int kpw_env = [] {  // tuning knob for the sweeps, not an API: 0 = always 4 keys, 1 = always 16
    const char* e = getenv("OMX_DECODE_KPW");
    return e ? atoi(e) : -1;
  }();
  const bool one_wave = kpw_env >= 0 ? kpw_env == 1 : pairs * p.num_splits <= sms;
  switch (a.q->dtype) {
    case OMX_FLOAT32: dd::launch_simt_f32(p, stream, Gt, grid, one_wave, want_cluster); break;
    case OMX_BFLOAT16: dd::launch_simt_bf16(p, stream, Gt, grid, one_wave, want_cluster); break;
    default: dd::launch_simt_f16(p, stream, Gt, grid, one_wave, want_cluster); break;
  }
  if (masked) masked_rows_fixup(a, p.dead, stream);
  if (f.ll) ll_second_kernel();  // (the CUDA-core kernel keeps the exchange as a kernel of its own)
}

#endif  // OMX_DECODE_PART == 0

}  // namespace omx
