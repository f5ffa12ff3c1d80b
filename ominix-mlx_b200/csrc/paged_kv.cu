// paged_kv.cu -- paged KV cache: page pool + block table + per-sequence lengths.
//
// north_star: "Decode attention ... fuses RoPE and the paged KV append".  The reference has no paged cache
// (mlx-rs-core/src/cache.rs:92-195 concatenates a fresh zero block every 256 tokens, an O(S) copy); what it
// fixes is the CONTRACT -- `KeyValueCache::{offset, update_and_fetch, reset}`, rows appended bit-for-bit,
// fetched [B, Hkv, offset, D] views -- and that contract is what this cache keeps, over a layout that a
// serving host needs on a 180 GB part:
//
//   pool K / pool V : [n_pages][Hkv][64][D], one allocation each, sized once (no reallocation, no copy on
//                     growth: a sequence grows by taking a page id from the free list)
//   block table     : int32 [B][max_pages] (device + host mirror): page of rows [64 t, 64 t + 64) of sequence b
//   lens            : int32 [2][B] (device, double-buffered by step parity) + host mirror: rows stored per
//                     sequence; sequences advance independently (ragged batches, slots released and reused)
//
// One page = 64 rows = one pipeline stage of the decode kernel = one TMA box pair per tensor, so the fused
// decode step (decode.cu, `paged` mode) streams pages exactly like rows of a contiguous cache: the producer
// lane reads the page id from the block table and issues the same four cp.async.bulk.tensor loads.
// update_and_fetch keeps the reference's return value available: the [B, Hkv, offset, D] views are
// MATERIALISED (gathered into a cache-owned contiguous buffer) only when the caller asks for them.
#include <algorithm>
#include <vector>

#include "omx_common.cuh"
#include "omx_internal.h"

namespace omx {

constexpr int kPageRows = 64;

struct PagedKVImpl {
  int B = 0, H = 0, Dk = 0, Dv = 0, dtype = 0;
  int64_t n_pages = 0;
  int max_pages = 0;  // per sequence
  void *kpool = nullptr, *vpool = nullptr;
  int* bt_dev = nullptr;    // [B][max_pages]
  int* lens_dev = nullptr;  // [2][B]
  int parity = 0;           // lens_dev[parity] is current
  std::vector<int> bt_host, lens_host, free_list;
  std::vector<int> pages_of;  // pages held per slot (== ceil(reserved rows / 64))
  int* stage = nullptr;       // pinned staging for table / length uploads: [B * max_pages + 2 * B]
  cudaEvent_t staged = nullptr;  // last upload that read `stage`
  bool stage_busy = false;
  // materialised views (on request)
  void *kflat = nullptr, *vflat = nullptr;
  int64_t flat_rows = 0;
  // split-K scratch owned by the cache (stable address under CUDA-graph capture)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  cudaStream_t last_stream = nullptr;
};

namespace {

// rows [r0, r0 + n) of every (b, h) <-> pages.  One thread per 16 bytes.  dir 0: src (strided [B,H,n,D]) ->
// pages at sequence rows start[b] + r; dir 1: pages -> dst (contiguous [B,H,rows,D]) for rows < len[b], zero beyond.
template <int DIR>
__global__ void paged_copy_kernel(uint4* __restrict__ pool, const int* __restrict__ bt, int bt_stride,
                                  const int* __restrict__ start, uint4* __restrict__ flat, int64_t fs0, int64_t fs1,
                                  int64_t fs2, int B, int H, int n, int vec_per_row, int slot0) {
  const int64_t total = (int64_t)B * H * n * vec_per_row;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % vec_per_row);
    int64_t r = idx / vec_per_row;
    const int row = (int)(r % n);
    r /= n;
    const int h = (int)(r % H);
    const int b = (int)(r / H);
    const int slot = slot0 + b;
    if (DIR == 0) {
      const int seq_row = start[slot] + row;
      const int64_t page = bt[(int64_t)slot * bt_stride + (seq_row >> 6)];
      pool[((page * H + h) * kPageRows + (seq_row & 63)) * vec_per_row + c] = flat[b * fs0 + h * fs1 + row * fs2 + c];
    } else {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (row < start[slot]) {  // start == length here
        const int64_t page = bt[(int64_t)slot * bt_stride + (row >> 6)];
        v = pool[((page * H + h) * kPageRows + (row & 63)) * vec_per_row + c];
      }
      flat[b * fs0 + h * fs1 + row * fs2 + c] = v;
    }
  }
}

void launch_paged_copy(int dir, void* pool, const int* bt, int bt_stride, const int* start, void* flat, int64_t fs0,
                       int64_t fs1, int64_t fs2, int B, int H, int n, int vec_per_row, int slot0,
                       cudaStream_t stream) {
  const int64_t total = (int64_t)B * H * n * vec_per_row;
  if (total == 0) return;
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, 148 * 16);
  if (dir == 0)
    paged_copy_kernel<0><<<blocks, threads, 0, stream>>>((uint4*)pool, bt, bt_stride, start, (uint4*)flat, fs0, fs1,
                                                         fs2, B, H, n, vec_per_row, slot0);
  else
    paged_copy_kernel<1><<<blocks, threads, 0, stream>>>((uint4*)pool, bt, bt_stride, start, (uint4*)flat, fs0, fs1,
                                                         fs2, B, H, n, vec_per_row, slot0);
  count_launch();
  OMX_CUDA(cudaGetLastError());
}

// lens[cur] is the truth; lens[other] is scratch that the next fused step overwrites for every active sequence
// and that stays -1 for released slots.  absolute: both buffers = value; else cur += delta for active slots.
__global__ void paged_set_lens_kernel(int* cur, int* other, int slot0, int n_slots, int value, int absolute) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  if (absolute) {
    cur[slot0 + i] = value;
    other[slot0 + i] = value;
  } else if (cur[slot0 + i] >= 0) {
    cur[slot0 + i] += value;
  }
}

void wait_stage(PagedKVImpl* c) {
  if (c->stage_busy) {
    OMX_CUDA(cudaEventSynchronize(c->staged));
    c->stage_busy = false;
  }
}

// make sure sequence `slot` owns pages for `rows` rows; uploads the new table entries
void ensure_pages(PagedKVImpl* c, int slot, int64_t rows, cudaStream_t stream, bool upload = true) {
  const int need = (int)((rows + kPageRows - 1) / kPageRows);
  OMX_CHECK(need <= c->max_pages, "[PagedKVCache] sequence %d needs %d pages; the block table holds %d per sequence "
            "(max_pages_per_seq at creation)", slot, need, c->max_pages);
  const int have = c->pages_of[slot];
  if (need <= have) return;
  OMX_CHECK((int)c->free_list.size() >= need - have, "[PagedKVCache] page pool exhausted: %d pages needed, %zu free "
            "of %lld (n_pages at creation)", need - have, c->free_list.size(), (long long)c->n_pages);
  for (int t = have; t < need; ++t) {
    c->bt_host[(size_t)slot * c->max_pages + t] = c->free_list.back();
    c->free_list.pop_back();
  }
  c->pages_of[slot] = need;
  if (upload) {
    // entries [have, need) of this slot's row: stage -> device (the stage is reused: wait for its last reader)
    wait_stage(c);
    const size_t off = (size_t)slot * c->max_pages + have;
    std::copy(c->bt_host.begin() + off, c->bt_host.begin() + off + (need - have), c->stage);
    OMX_CUDA(cudaMemcpyAsync(c->bt_dev + off, c->stage, sizeof(int) * (need - have), cudaMemcpyHostToDevice, stream));
    OMX_CUDA(cudaEventRecord(c->staged, stream));
    c->stage_busy = true;
  }
}

void upload_table(PagedKVImpl* c, cudaStream_t stream) {
  wait_stage(c);
  std::copy(c->bt_host.begin(), c->bt_host.end(), c->stage);
  OMX_CUDA(cudaMemcpyAsync(c->bt_dev, c->stage, sizeof(int) * c->bt_host.size(), cudaMemcpyHostToDevice, stream));
  OMX_CUDA(cudaEventRecord(c->staged, stream));
  c->stage_busy = true;
}

}  // namespace

PagedKVImpl* paged_create(int B, int H, int Dk, int Dv, int dtype, int64_t n_pages, int max_pages_per_seq) {
  OMX_CHECK(B >= 1 && H >= 1 && Dk >= 1 && Dv >= 1, "[PagedKVCache] bad geometry [%d, %d, *, %d/%d]", B, H, Dk, Dv);
  OMX_CHECK(is_float_dtype(dtype), "[PagedKVCache] dtype must be floating point");
  OMX_CHECK(n_pages >= 1 && max_pages_per_seq >= 1, "[PagedKVCache] n_pages and max_pages_per_seq must be positive");
  const size_t es = dtype_size(dtype);
  OMX_CHECK((Dk * es) % 16 == 0 && (Dv * es) % 16 == 0, "[PagedKVCache] head_dim rows must be multiples of 16 bytes");
  auto* c = new PagedKVImpl();
  c->B = B; c->H = H; c->Dk = Dk; c->Dv = Dv; c->dtype = dtype;
  c->n_pages = n_pages;
  c->max_pages = max_pages_per_seq;
  try {
    OMX_CUDA(cudaMalloc(&c->kpool, (size_t)n_pages * H * kPageRows * Dk * es));
    OMX_CUDA(cudaMalloc(&c->vpool, (size_t)n_pages * H * kPageRows * Dv * es));
    // never-written rows read as +0.0 (a partial last page is masked by the kernels, but must hold finite data)
    OMX_CUDA(cudaMemset(c->kpool, 0, (size_t)n_pages * H * kPageRows * Dk * es));
    OMX_CUDA(cudaMemset(c->vpool, 0, (size_t)n_pages * H * kPageRows * Dv * es));
    OMX_CUDA(cudaMalloc(&c->bt_dev, sizeof(int) * (size_t)B * max_pages_per_seq));
    OMX_CUDA(cudaMemset(c->bt_dev, 0, sizeof(int) * (size_t)B * max_pages_per_seq));
    OMX_CUDA(cudaMalloc(&c->lens_dev, sizeof(int) * 2 * B));
    OMX_CUDA(cudaMemset(c->lens_dev, 0, sizeof(int) * 2 * B));
    OMX_CUDA(cudaMallocHost(&c->stage, sizeof(int) * ((size_t)B * max_pages_per_seq + 2 * B)));
    OMX_CUDA(cudaEventCreateWithFlags(&c->staged, cudaEventDisableTiming));
  } catch (...) {
    paged_destroy(c);
    throw;
  }
  c->bt_host.assign((size_t)B * max_pages_per_seq, 0);
  c->lens_host.assign(B, 0);
  c->pages_of.assign(B, 0);
  c->free_list.resize(n_pages);
  for (int64_t i = 0; i < n_pages; ++i) c->free_list[i] = (int)(n_pages - 1 - i);  // page 0 is handed out first
  return c;
}

void paged_destroy(PagedKVImpl* c) {
  if (!c) return;
  cudaDeviceSynchronize();
  for (void* p : {c->kpool, c->vpool, (void*)c->bt_dev, (void*)c->lens_dev, c->kflat, c->vflat, c->scratch})
    if (p) cudaFree(p);
  if (c->stage) cudaFreeHost(c->stage);
  if (c->staged) cudaEventDestroy(c->staged);
  delete c;
}

int paged_offset(const PagedKVImpl* c) { return *std::max_element(c->lens_host.begin(), c->lens_host.end()); }
const int* paged_lengths(const PagedKVImpl* c) { return c->lens_host.data(); }
int paged_batch(const PagedKVImpl* c) { return c->B; }
int64_t paged_free_pages(const PagedKVImpl* c) { return (int64_t)c->free_list.size(); }

void paged_shape(const PagedKVImpl* c, int* B, int* H, int* Dk, int* Dv, int* dtype, int64_t* n_pages, int* max_pages) {
  *B = c->B; *H = c->H; *Dk = c->Dk; *Dv = c->Dv; *dtype = c->dtype; *n_pages = c->n_pages; *max_pages = c->max_pages;
}

// slot < 0: every sequence.  Pages go back to the free list, the length becomes 0 (reset()) or -1 (release:
// the fused decode step skips the slot).
void paged_reset(PagedKVImpl* c, int slot, bool deactivate, cudaStream_t stream) {
  OMX_CHECK(slot < c->B, "[PagedKVCache] slot %d out of range (batch %d)", slot, c->B);
  // pages may still be read by launches in flight on other streams: the caller orders those; launches on
  // `stream` are ordered by the length upload below
  const int s0 = slot < 0 ? 0 : slot, s1 = slot < 0 ? c->B : slot + 1;
  for (int s = s0; s < s1; ++s) {
    for (int t = 0; t < c->pages_of[s]; ++t) c->free_list.push_back(c->bt_host[(size_t)s * c->max_pages + t]);
    c->pages_of[s] = 0;
    c->lens_host[s] = deactivate ? -1 : 0;
  }
  paged_set_lens_kernel<<<(s1 - s0 + 127) / 128, 128, 0, stream>>>(c->lens_dev + c->parity * c->B,
                                                                   c->lens_dev + (c->parity ^ 1) * c->B, s0, s1 - s0,
                                                                   deactivate ? -1 : 0, 1);
  count_launch();
  OMX_CUDA(cudaGetLastError());
  c->last_stream = stream;
}

// Pre-assign pages for `rows_ahead` more rows of every active sequence: the next rows_ahead fused decode steps
// then need no host-side allocation (and can be captured into a CUDA graph, two steps per capture).
void paged_reserve(PagedKVImpl* c, int rows_ahead, cudaStream_t stream) {
  OMX_CHECK(rows_ahead >= 0, "[PagedKVCache] reserve: negative row count");
  bool any = false;
  for (int s = 0; s < c->B; ++s) {
    if (c->lens_host[s] < 0) continue;
    const int before = c->pages_of[s];
    ensure_pages(c, s, (int64_t)c->lens_host[s] + rows_ahead, stream, /*upload=*/false);
    any = any || c->pages_of[s] != before;
  }
  if (any) upload_table(c, stream);
  c->last_stream = stream;
}

// Append n rows to sequences [slot0, slot0 + keys.B) at each sequence's own length (cache.rs:183-188 per
// sequence).  keys / values: [Bs, Hkv, n, D] strided views, feature axis contiguous.
void paged_append(PagedKVImpl* c, int slot0, const omx_array* keys, const omx_array* values, cudaStream_t stream) {
  OMX_CHECK(keys && values && keys->ndim == 4 && values->ndim == 4,
            "[PagedKVCache] keys and values must be 4-dimensional [B, n_kv_heads, n, head_dim]");
  for (int i = 0; i < 3; ++i)
    OMX_CHECK(keys->shape[i] == values->shape[i], "[PagedKVCache] keys/values shape mismatch on axis %d", i);
  const int Bs = (int)keys->shape[0], n = (int)keys->shape[2];
  OMX_CHECK(slot0 >= 0 && slot0 + Bs <= c->B && keys->shape[1] == c->H && keys->shape[3] == c->Dk &&
                values->shape[3] == c->Dv,
            "[PagedKVCache] update shape [%lld,%lld,%lld,%lld] at slot %d does not match the cache [%d,%d,*,%d]",
            (long long)keys->shape[0], (long long)keys->shape[1], (long long)keys->shape[2], (long long)keys->shape[3],
            slot0, c->B, c->H, c->Dk);
  OMX_CHECK(keys->dtype == c->dtype && values->dtype == c->dtype, "[PagedKVCache] update dtype differs from the cache dtype");
  c->last_stream = stream;
  if (n == 0 || Bs == 0) return;
  const size_t es = dtype_size(c->dtype);
  const int64_t v16 = (int64_t)(16 / es);
  for (const omx_array* t : {keys, values}) {
    OMX_CHECK((t->strides[3] == 1 || t->shape[3] == 1) && aligned16(t->data) && t->strides[0] % v16 == 0 &&
                  t->strides[1] % v16 == 0 && t->strides[2] % v16 == 0,
              "[PagedKVCache] keys / values need a contiguous feature axis and 16-byte aligned strides");
  }
  for (int b = 0; b < Bs; ++b) {
    OMX_CHECK(c->lens_host[slot0 + b] >= 0, "[PagedKVCache] slot %d was released; reset it before appending", slot0 + b);
    ensure_pages(c, slot0 + b, (int64_t)c->lens_host[slot0 + b] + n, stream);
  }
  const int* lens_cur = c->lens_dev + c->parity * c->B;
  launch_paged_copy(0, c->kpool, c->bt_dev, c->max_pages, lens_cur, keys->data, keys->strides[0] / v16,
                    keys->strides[1] / v16, keys->strides[2] / v16, Bs, c->H, n, (int)(c->Dk / v16), slot0, stream);
  launch_paged_copy(0, c->vpool, c->bt_dev, c->max_pages, lens_cur, values->data, values->strides[0] / v16,
                    values->strides[1] / v16, values->strides[2] / v16, Bs, c->H, n, (int)(c->Dv / v16), slot0, stream);
  paged_set_lens_kernel<<<(Bs + 127) / 128, 128, 0, stream>>>(c->lens_dev + c->parity * c->B,
                                                              c->lens_dev + (c->parity ^ 1) * c->B, slot0, Bs, n, 0);
  count_launch();
  OMX_CUDA(cudaGetLastError());
  for (int b = 0; b < Bs; ++b) c->lens_host[slot0 + b] += n;
}

// The reference's return value of update_and_fetch: [B, Hkv, offset, D] views (cache.rs:190-193), gathered into a
// cache-owned contiguous buffer.  offset = the longest sequence; rows past a shorter sequence's length read +0.0.
void paged_materialize(PagedKVImpl* c, omx_array* keys_out, omx_array* values_out, cudaStream_t stream) {
  const int rows = std::max(paged_offset(c), 0);
  const size_t es = dtype_size(c->dtype);
  if (rows > c->flat_rows) {
    const int64_t cap = std::max<int64_t>(((int64_t)rows + 255) / 256 * 256, 2 * c->flat_rows);
    if (c->kflat) OMX_CUDA(cudaFreeAsync(c->kflat, stream));
    if (c->vflat) OMX_CUDA(cudaFreeAsync(c->vflat, stream));
    OMX_CUDA(cudaMallocAsync(&c->kflat, (size_t)c->B * c->H * cap * c->Dk * es, stream));
    OMX_CUDA(cudaMallocAsync(&c->vflat, (size_t)c->B * c->H * cap * c->Dv * es, stream));
    c->flat_rows = cap;
  }
  const int64_t v16 = (int64_t)(16 / es);
  const int* lens_cur = c->lens_dev + c->parity * c->B;
  auto fill = [&](omx_array* out, void* flat, int D, void* pool) {
    if (rows > 0)
      launch_paged_copy(1, pool, c->bt_dev, c->max_pages, lens_cur, flat, (int64_t)c->H * c->flat_rows * D / v16,
                        c->flat_rows * D / v16, D / v16, c->B, c->H, rows, (int)(D / v16), 0, stream);
    if (!out) return;
    out->data = flat; out->dtype = c->dtype; out->ndim = 4;
    out->shape[0] = c->B; out->shape[1] = c->H; out->shape[2] = rows; out->shape[3] = D;
    out->strides[0] = (int64_t)c->H * c->flat_rows * D; out->strides[1] = c->flat_rows * D;
    out->strides[2] = D; out->strides[3] = 1;
  };
  fill(keys_out, c->kflat, c->Dk, c->kpool);
  fill(values_out, c->vflat, c->Dv, c->vpool);
  c->last_stream = stream;
}

// ---- the fused decode step over the pages: everything decode_attention needs
void paged_begin_step(PagedKVImpl* c, int n_q_heads, omx_array* kpool_view, omx_array* vpool_view, PagedRef* ref,
                      void** scratch, size_t* scratch_bytes, int* max_len_after, int* table_rows,
                      cudaStream_t stream) {
  // pages for the row each active sequence is about to receive (host-side only when a sequence crosses a page
  // boundary and nothing was reserved)
  int mx = 0, held = 0;
  for (int s = 0; s < c->B; ++s) {
    if (c->lens_host[s] < 0) continue;
    ensure_pages(c, s, (int64_t)c->lens_host[s] + 1, stream);
    mx = std::max(mx, c->lens_host[s] + 1);
    held = std::max(held, c->pages_of[s] * kPageRows);
  }
  *max_len_after = mx;
  // rope rows a (replayed) launch may look up: every position the sequences hold pages for
  *table_rows = std::max(mx, held);
  // split-K partials of the decode kernels, or the prologue's q' rows + the mma.sync kernel's partials (omx_api.cu)
  const size_t qbytes = ((size_t)c->B * n_q_heads * c->Dk * dtype_size(c->dtype) + 255) & ~(size_t)255;
  // (only geometries that can take the mma.sync route pay for its scratch: 16-bit, not the 128 / 128 heads of the TMA kernel)
  const bool mma_route = (c->dtype == OMX_BFLOAT16 || c->dtype == OMX_FLOAT16) && !(c->Dk == 128 && c->Dv == 128);
  const size_t need = std::max(decode_graph_scratch_bytes(c->B, c->H, n_q_heads, c->Dk, c->dtype, c->max_pages * kPageRows),
                               mma_route ? qbytes + sdpa_mma_graph_scratch_bytes(c->B, c->H, n_q_heads, c->Dv) : (size_t)0);
  if (need > c->scratch_bytes) {
    if (c->scratch) OMX_CUDA(cudaFreeAsync(c->scratch, stream));
    OMX_CUDA(cudaMallocAsync(&c->scratch, need, stream));
    OMX_CUDA(cudaMemsetAsync(c->scratch, 0, need, stream));
    c->scratch_bytes = need;
  }
  *scratch = c->scratch;
  *scratch_bytes = c->scratch_bytes;
  auto pool_view = [&](omx_array* v, void* pool, int D) {
    v->data = pool; v->dtype = c->dtype; v->ndim = 4;
    v->shape[0] = c->B; v->shape[1] = c->H; v->shape[2] = mx; v->shape[3] = D;
    v->strides[0] = (int64_t)c->H * kPageRows * D;  // PAGE stride (see DecodeParams::paged)
    v->strides[1] = (int64_t)kPageRows * D;
    v->strides[2] = D; v->strides[3] = 1;
  };
  pool_view(kpool_view, c->kpool, c->Dk);
  pool_view(vpool_view, c->vpool, c->Dv);
  ref->block_table = c->bt_dev;
  ref->bt_stride = c->max_pages;
  ref->lens_in = c->lens_dev + c->parity * c->B;
  ref->lens_out = c->lens_dev + (c->parity ^ 1) * c->B;
  ref->n_pages = c->n_pages;
  c->last_stream = stream;
}

// after the launch: the other length buffer is current; inactive slots keep -1 in both
void paged_end_step(PagedKVImpl* c) {
  c->parity ^= 1;
  for (int s = 0; s < c->B; ++s)
    if (c->lens_host[s] >= 0) ++c->lens_host[s];
}

// Drop the last n rows of every active sequence (bench / speculative decoding rewind): lengths only.
void paged_trim(PagedKVImpl* c, int n, cudaStream_t stream) {
  OMX_CHECK(n >= 0, "[PagedKVCache] trim by a negative count");
  int mn = 1 << 30;
  for (int s = 0; s < c->B; ++s)
    if (c->lens_host[s] >= 0) mn = std::min(mn, c->lens_host[s]);
  if (mn == (1 << 30) || n == 0) return;
  OMX_CHECK(n <= mn, "[PagedKVCache] trim(%d) exceeds the shortest active sequence (%d rows)", n, mn);
  for (int s = 0; s < c->B; ++s)
    if (c->lens_host[s] >= 0) c->lens_host[s] -= n;
  paged_set_lens_kernel<<<(c->B + 127) / 128, 128, 0, stream>>>(c->lens_dev + c->parity * c->B,
                                                                c->lens_dev + (c->parity ^ 1) * c->B, 0, c->B, -n, 0);
  count_launch();
  OMX_CUDA(cudaGetLastError());
  c->last_stream = stream;
}

// Host mirror <- device lengths (synchronises `stream`).  Needed after CUDA-graph capture / replays of the fused
// step, which advance the device lengths without the host seeing it.
void paged_sync_lengths(PagedKVImpl* c, cudaStream_t stream) {
  wait_stage(c);
  int* st = c->stage + (size_t)c->B * c->max_pages;
  OMX_CUDA(cudaMemcpyAsync(st, c->lens_dev + c->parity * c->B, sizeof(int) * c->B, cudaMemcpyDeviceToHost, stream));
  OMX_CUDA(cudaStreamSynchronize(stream));
  for (int s = 0; s < c->B; ++s) {
    OMX_CHECK(st[s] <= c->pages_of[s] * kPageRows, "[PagedKVCache] sequence %d ran to %d rows but holds pages for %d: "
              "replays went past the reserved pages", s, st[s], c->pages_of[s] * kPageRows);
    c->lens_host[s] = st[s];
  }
}

void paged_pool_ptrs(const PagedKVImpl* c, void** kpool, void** vpool, const int** block_table) {
  *kpool = c->kpool; *vpool = c->vpool; *block_table = c->bt_host.data();
}

}  // namespace omx
