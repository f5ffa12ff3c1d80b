// kv_cache.cu -- device-resident KV caches with the reference's growth contract.
//
// Replaces the Rust logic of mlx-rs-core/src/cache.rs:
//   KVCache (step-allocated, :92-195) and ConcatKeyValueCache (:45-85),
// which the reference builds from mlx_zeros / mlx_concatenate_axis / mlx_slice_update /
// mlx_slice (mlx-c/mlx/c/ops.h:210,960-989,1220).
//
// HBM layout: one allocation per tensor, [B, Hkv, phys_rows, D] row-major; the LOGICAL
// capacity `cap` (= the reference's keys.shape[2]) follows cache.rs:141-181 exactly, the
// PHYSICAL row count only ever doubles (or is pre-sized with omx_kv_cache_reserve), so a
// decode loop re-allocates O(log S) times instead of every `step` tokens.  Fetched views
// therefore have head stride phys_rows*D (the reference's views have cap*D; both are plain
// strided views to the consumer).  Rows [offset, cap) hold +0.0 or stale rows after reset(),
// exactly as in the reference.
#include <algorithm>

#include "omx_common.cuh"
#include "omx_internal.h"

namespace omx {

namespace {

template <typename U>
__global__ void copy4d_kernel(U* __restrict__ dst, const U* __restrict__ src, int64_t n0, int64_t n1,
                              int64_t n2, int64_t n3, int64_t d0, int64_t d1, int64_t d2, int64_t d3,
                              int64_t s0, int64_t s1, int64_t s2, int64_t s3) {
  const int64_t total = n0 * n1 * n2 * n3;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i3 = idx % n3;
    int64_t r = idx / n3;
    const int64_t i2 = r % n2;
    r /= n2;
    const int64_t i1 = r % n1;
    const int64_t i0 = r / n1;
    dst[i0 * d0 + i1 * d1 + i2 * d2 + i3 * d3] = src[i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3];
  }
}

template <typename U>
void launch_copy(void* dst, const void* src, const int64_t n[4], const int64_t d[4],
                 const int64_t s[4], cudaStream_t stream) {
  const int64_t total = n[0] * n[1] * n[2] * n[3];
  if (total == 0) return;
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, 148 * 32);
  copy4d_kernel<U><<<blocks, threads, 0, stream>>>((U*)dst, (const U*)src, n[0], n[1], n[2], n[3], d[0],
                                                   d[1], d[2], d[3], s[0], s[1], s[2], s[3]);
  count_launch();
  OMX_CUDA(cudaGetLastError());
}

}  // namespace

void copy4d(const omx_array* dst, const omx_array* src, cudaStream_t stream) {
  OMX_CHECK(dst->ndim == 4 && src->ndim == 4 && dst->dtype == src->dtype, "[copy4d] bad arguments");
  const size_t es = dtype_size(src->dtype);
  int64_t n[4], d[4], s[4];
  for (int i = 0; i < 4; ++i) {
    OMX_CHECK(dst->shape[i] == src->shape[i], "[copy4d] shape mismatch");
    n[i] = src->shape[i];
    d[i] = dst->strides[i];
    s[i] = src->strides[i];
  }
  // 16-byte path: innermost axis contiguous and everything 16 B aligned
  const int64_t v = (int64_t)(16 / es);
  bool vec = d[3] == 1 && s[3] == 1 && n[3] % v == 0 && aligned16(dst->data) && aligned16(src->data);
  for (int i = 0; i < 3; ++i) vec = vec && d[i] % v == 0 && s[i] % v == 0;
  if (vec) {
    int64_t nv[4] = {n[0], n[1], n[2], n[3] / v};
    int64_t dv[4] = {d[0] / v, d[1] / v, d[2] / v, 1};
    int64_t sv[4] = {s[0] / v, s[1] / v, s[2] / v, 1};
    launch_copy<uint4>(dst->data, src->data, nv, dv, sv, stream);
  } else if (es == 4) {
    launch_copy<uint32_t>(dst->data, src->data, n, d, s, stream);
  } else if (es == 2) {
    launch_copy<uint16_t>(dst->data, src->data, n, d, s, stream);
  } else {
    launch_copy<uint8_t>(dst->data, src->data, n, d, s, stream);
  }
}

struct KVBuf {
  void* p = nullptr;
  int D = 0;
  int dtype = 0;
  int64_t phys = 0;  // physical rows
  // The buffer a growth step replaced.  The reference hands out refcounted arrays that outlive a growth; here
  // fetched views are borrowed, so the previous buffer stays allocated (and readable, e.g. by a consumer still
  // running on another stream) until the NEXT growth or the cache's destruction: views survive one growth step.
  void* retired = nullptr;
};

struct KVCacheImpl {
  bool concat = false;
  int step = 256;
  int offset = 0;
  bool has = false;
  int64_t cap = 0;  // logical rows (reference's keys.shape[2])
  int B = 0, H = 0;
  int64_t reserve_rows = 0;
  KVBuf k, v;
  cudaStream_t last_stream = nullptr;
  // Rows [0, stable_now) were written by calls that precede the cache's most recent writing call, i.e. not by
  // the kernel that may still be running when the next launch on the stream is dispatched early (programmatic
  // dependent launch, decode.cu): the fused decode step may request them before its dependency wait.
  // stable_next = rows in place before the most recent write (becomes stable_now at the next update);
  // a call that grows / moves / zero-fills the buffer starts from 0.
  int stable_next = 0, stable_now = 0;
  // graph mode
  int64_t graph_rows = 0;  // rows pinned by prepare_graph (0: not prepared)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
};

KVCacheImpl* kv_cache_create(int step, bool concat) {
  OMX_CHECK(concat || step > 0, "[KVCache] step must be positive, got %d", step);
  auto* c = new KVCacheImpl();
  c->concat = concat;
  c->step = step;
  return c;
}

void kv_cache_destroy(KVCacheImpl* c) {
  if (!c) return;
  for (void* p : {c->k.p, c->v.p, c->k.retired, c->v.retired})
    if (p) cudaFreeAsync(p, c->last_stream);
  if (c->scratch) cudaFreeAsync(c->scratch, c->last_stream);
  delete c;
}

int kv_cache_offset(const KVCacheImpl* c) { return c->offset; }
int kv_cache_stable_rows(const KVCacheImpl* c) { return c->stable_now; }
bool kv_cache_is_concat(const KVCacheImpl* c) { return c->concat; }

void kv_cache_reset(KVCacheImpl* c) {
  if (!c->concat) c->offset = 0;  // cache.rs:130-132; the concat cache keeps the trait's no-op
}

int kv_cache_trim(KVCacheImpl* c, int n) {
  OMX_CHECK(!c->concat, "[KVCache] trim is not defined for ConcatKeyValueCache");
  const int t = std::max(0, std::min(n, c->offset));
  c->offset -= t;
  c->stable_next = std::min(c->stable_next, c->offset);  // trimmed rows are rewritten by later steps
  c->stable_now = std::min(c->stable_now, c->offset);
  return t;
}

void kv_cache_reserve(KVCacheImpl* c, int rows) {
  OMX_CHECK(rows >= 0, "[KVCache] reserve rows must be >= 0");
  c->reserve_rows = rows;
}

namespace {

// Make `b` hold `new_cap` logical rows: rows [0, keep) preserved, rows [keep, new_cap) zeroed.
void regrow(KVCacheImpl* c, KVBuf& b, int64_t keep, int64_t new_cap, bool zero_new,
            cudaStream_t stream) {
  const size_t es = dtype_size(b.dtype);
  const size_t row = (size_t)b.D * es;
  const size_t heads = (size_t)c->B * c->H;
  if (new_cap > b.phys) {
    int64_t phys = std::max<int64_t>(new_cap, std::max<int64_t>(c->reserve_rows, 2 * b.phys));
    void* np = nullptr;
    OMX_CUDA(cudaMallocAsync(&np, heads * (size_t)phys * row, stream));
    if (b.p && keep > 0) {
      OMX_CUDA(cudaMemcpy2DAsync(np, (size_t)phys * row, b.p, (size_t)b.phys * row, (size_t)keep * row,
                                 heads, cudaMemcpyDeviceToDevice, stream));
    }
    if (b.retired) OMX_CUDA(cudaFreeAsync(b.retired, stream));
    b.retired = b.p;
    b.p = np;
    b.phys = phys;
    c->graph_rows = 0;  // the buffer moved: launches captured against the old address are stale
  }
  if (zero_new && new_cap > keep && row > 0 && heads > 0) {
    OMX_CUDA(cudaMemset2DAsync((char*)b.p + (size_t)keep * row, (size_t)b.phys * row, 0,
                               (size_t)(new_cap - keep) * row, heads, stream));
  }
}

void fill_view(const KVCacheImpl* c, const KVBuf& b, int64_t rows, omx_array* out) {
  if (!out) return;
  out->data = b.p;
  out->dtype = b.dtype;
  out->ndim = 4;
  out->shape[0] = c->B;
  out->shape[1] = c->H;
  out->shape[2] = rows;
  out->shape[3] = b.D;
  out->strides[0] = (int64_t)c->H * b.phys * b.D;
  out->strides[1] = b.phys * b.D;
  out->strides[2] = b.D;
  out->strides[3] = 1;
}

}  // namespace

void kv_cache_update(KVCacheImpl* c, const omx_array* keys, const omx_array* values,
                     omx_array* keys_out, omx_array* values_out, bool skip_copy,
                     cudaStream_t stream) {
  OMX_CHECK(keys && values && keys->ndim == 4 && values->ndim == 4,
            "[KVCache] keys and values must be 4-dimensional [B, n_kv_heads, n, head_dim]");
  OMX_CHECK(is_float_dtype(keys->dtype) && is_float_dtype(values->dtype),
            "[KVCache] keys/values must be floating point");
  for (int i = 0; i < 3; ++i)
    OMX_CHECK(keys->shape[i] == values->shape[i], "[KVCache] keys/values shape mismatch on axis %d", i);
  const int prev = c->offset;
  const int n = (int)keys->shape[2];
  // rows the launch of THIS call may read ahead of its dependency wait (see KVCacheImpl::stable_now); a change
  // of stream is ordered by the caller with events, which says nothing about the new stream's previous kernel
  c->stable_now = (c->concat || stream != c->last_stream) ? 0 : std::min(c->stable_next, prev);
  c->stable_next = prev;
  c->last_stream = stream;

  if (c->has) {
    OMX_CHECK(keys->shape[0] == c->B && keys->shape[1] == c->H && keys->shape[3] == c->k.D &&
                  values->shape[3] == c->v.D,
              "[KVCache] update shape [%lld,%lld,%lld,%lld] does not match the cache [%d,%d,*,%d]",
              (long long)keys->shape[0], (long long)keys->shape[1], (long long)keys->shape[2],
              (long long)keys->shape[3], c->B, c->H, c->k.D);
    OMX_CHECK(keys->dtype == c->k.dtype && values->dtype == c->v.dtype,
              "[KVCache] update dtype differs from the cache dtype");
  }

  if (c->concat) {
    // cache.rs:66-84: keys = concat([old, new], -2); offset = keys.shape[-2]
    if (!c->has) {
      c->B = (int)keys->shape[0];
      c->H = (int)keys->shape[1];
      c->k.D = (int)keys->shape[3];
      c->v.D = (int)values->shape[3];
      c->k.dtype = keys->dtype;
      c->v.dtype = values->dtype;
      c->has = true;
    }
    const int64_t new_cap = c->cap + n;
    regrow(c, c->k, c->cap, new_cap, false, stream);
    regrow(c, c->v, c->cap, new_cap, false, stream);
    const int64_t at = c->cap;
    c->cap = new_cap;
    c->offset = (int)new_cap;
    if (n > 0 && !skip_copy) {
      omx_array dk, dv;
      fill_view(c, c->k, n, &dk);
      fill_view(c, c->v, n, &dv);
      dk.data = (char*)c->k.p + (size_t)at * c->k.D * dtype_size(c->k.dtype);
      dv.data = (char*)c->v.p + (size_t)at * c->v.D * dtype_size(c->v.dtype);
      copy4d(&dk, keys, stream);
      copy4d(&dv, values, stream);
    }
    fill_view(c, c->k, c->offset, keys_out);
    fill_view(c, c->v, c->offset, values_out);
    return;
  }

  // ---- KVCache::update_and_fetch, cache.rs:134-194 ----
  const bool needs_grow = !c->has || (int64_t)prev + n > c->cap;  // :141-144
  if (needs_grow) {
    const int n_steps = (c->step + n - 1) / c->step;  // :152
    const int64_t new_size = (int64_t)n_steps * c->step;
    int64_t keep = 0;
    if (!c->has) {
      c->B = (int)keys->shape[0];  // :147-150
      c->H = (int)keys->shape[1];
      c->k.D = (int)keys->shape[3];
      c->v.D = (int)values->shape[3];
      c->k.dtype = keys->dtype;  // :158-161
      c->v.dtype = values->dtype;
      c->has = true;
    } else {
      keep = (prev % c->step != 0) ? prev : c->cap;  // :165-172 (trim to prev, else keep whole)
    }
    const int64_t new_cap = keep + new_size;  // :173-174 concat([old, zeros])
    regrow(c, c->k, keep, new_cap, true, stream);
    regrow(c, c->v, keep, new_cap, true, stream);
    c->cap = new_cap;
    c->stable_now = 0;  // copies / zero fills of this very call precede the launch
  }
  c->offset = prev + n;  // :183
  if (n > 0 && !skip_copy) {  // :187-188 slice update
    omx_array dk, dv;
    fill_view(c, c->k, n, &dk);
    fill_view(c, c->v, n, &dv);
    dk.data = (char*)c->k.p + (size_t)prev * c->k.D * dtype_size(c->k.dtype);
    dv.data = (char*)c->v.p + (size_t)prev * c->v.D * dtype_size(c->v.dtype);
    copy4d(&dk, keys, stream);
    copy4d(&dv, values, stream);
  }
  fill_view(c, c->k, c->offset, keys_out);  // :190-193
  fill_view(c, c->v, c->offset, values_out);
}

void kv_cache_prepare_graph(KVCacheImpl* c, int max_rows, size_t scratch_bytes, cudaStream_t stream) {
  OMX_CHECK(!c->concat, "[KVCache] graph mode is not defined for ConcatKeyValueCache");
  OMX_CHECK(c->has, "[KVCache] prepare_graph needs a cache that has seen one update (shape and dtype come from it)");
  OMX_CHECK(max_rows >= c->offset + 1, "[KVCache] prepare_graph: max_rows %d leaves no room after offset %d", max_rows,
            c->offset);
  c->last_stream = stream;
  c->stable_next = c->stable_now = 0;  // the buffers may move
  // the logical capacity can run up to one step past the last position (cache.rs:152-174): pin that too
  const int64_t pin = ((int64_t)max_rows + c->step - 1) / c->step * c->step + c->step;
  c->reserve_rows = std::max<int64_t>(c->reserve_rows, pin);
  for (KVBuf* b : {&c->k, &c->v}) {
    regrow(c, *b, c->cap, std::max<int64_t>(c->cap, pin), false, stream);  // may move the buffer
    // physical rows past the logical capacity have never been exposed: zero them once so that a partial
    // tile read beyond the position sees finite data (fresh allocations are not zeroed by the driver)
    const size_t row = (size_t)b->D * dtype_size(b->dtype);
    const size_t heads = (size_t)c->B * c->H;
    const int64_t from = c->cap;
    if (b->phys > from && row && heads)
      OMX_CUDA(cudaMemset2DAsync((char*)b->p + (size_t)from * row, (size_t)b->phys * row, 0,
                                 (size_t)(b->phys - from) * row, heads, stream));
  }
  if (scratch_bytes > c->scratch_bytes) {
    if (c->scratch) OMX_CUDA(cudaFreeAsync(c->scratch, stream));
    OMX_CUDA(cudaMallocAsync(&c->scratch, scratch_bytes, stream));
    OMX_CUDA(cudaMemsetAsync(c->scratch, 0, scratch_bytes, stream));
    c->scratch_bytes = scratch_bytes;
  }
  c->graph_rows = max_rows;
}

bool kv_cache_graph_view(const KVCacheImpl* c, omx_array* k, omx_array* v, void** scratch, size_t* scratch_bytes,
                         int* max_rows) {
  if (!c->has || c->graph_rows <= 0 || c->k.phys < c->graph_rows || c->v.phys < c->graph_rows) return false;
  fill_view(c, c->k, c->graph_rows, k);
  fill_view(c, c->v, c->graph_rows, v);
  *scratch = c->scratch;
  *scratch_bytes = c->scratch_bytes;
  *max_rows = (int)c->graph_rows;
  return true;
}

void kv_cache_advance(KVCacheImpl* c, int n, cudaStream_t stream) {
  OMX_CHECK(!c->concat && c->has, "[KVCache] advance needs a non-empty step-allocated cache");
  OMX_CHECK(n >= 0, "[KVCache] advance by a negative count");
  OMX_CHECK(c->graph_rows > 0 && (int64_t)c->offset + n <= c->graph_rows,
            "[KVCache] advance: offset %d + %d rows exceeds the %lld rows pinned by prepare_graph (the launches "
            "past that point rewrote the last row)", c->offset, n, (long long)c->graph_rows);
  const int prev = c->offset;
  c->last_stream = stream;
  if ((int64_t)prev + n > c->cap) {  // cache.rs:141-181, with the rows the kernels already wrote kept
    const int n_steps = (c->step + n - 1) / c->step;
    const int64_t keep = (prev % c->step != 0) ? prev : c->cap;
    const int64_t new_cap = keep + (int64_t)n_steps * c->step;
    // the reference zero-fills [keep, new_cap) and then writes the n rows; here the rows are already in
    // place, so only the part after them is cleared (rows beyond the old capacity are zero since
    // prepare_graph unless a trim exposed stale ones)
    const int64_t from = std::max<int64_t>(keep, (int64_t)prev + n);
    for (KVBuf* b : {&c->k, &c->v}) {
      OMX_CHECK(new_cap <= b->phys, "[KVCache] advance: capacity %lld exceeds the pinned rows", (long long)new_cap);
      const size_t row = (size_t)b->D * dtype_size(b->dtype);
      const size_t heads = (size_t)c->B * c->H;
      const int64_t to = std::min<int64_t>(new_cap, b->phys);
      if (to > from && row && heads)
        OMX_CUDA(cudaMemset2DAsync((char*)b->p + (size_t)from * row, (size_t)b->phys * row, 0,
                                   (size_t)(to - from) * row, heads, stream));
    }
    c->cap = new_cap;
  }
  c->offset = prev + n;
}

KVCacheSnapshot kv_cache_snapshot(const KVCacheImpl* c) { return {c->offset, c->cap, c->has}; }

void kv_cache_rollback(KVCacheImpl* c, const KVCacheSnapshot& s, cudaStream_t stream) {
  if (!s.has && c->has) {
    // the failed call was the cache's first update: forget the shape / dtype it latched and its buffers
    for (KVBuf* b : {&c->k, &c->v}) {
      if (b->p) cudaFreeAsync(b->p, stream);
      if (b->retired) cudaFreeAsync(b->retired, stream);
      *b = KVBuf();
    }
    c->has = false;
    c->B = c->H = 0;
    c->graph_rows = 0;
  }
  // rows [offset, cap) that a growth zero-filled stay zero; they are past the offset, i.e. never fetched
  c->offset = s.offset;
  c->cap = s.cap;
  c->stable_next = std::min(c->stable_next, c->offset);
  c->stable_now = 0;
}

void kv_cache_shape(const KVCacheImpl* c, int* B, int* H, int* Dk, int* Dv, int* dtype) {
  OMX_CHECK(c->has, "[KVCache] cache is empty");
  *B = c->B; *H = c->H; *Dk = c->k.D; *Dv = c->v.D; *dtype = c->k.dtype;
}

void kv_cache_state(const KVCacheImpl* c, omx_array* kbuf, omx_array* vbuf) {
  OMX_CHECK(c->has, "[KVCache] cache is empty");
  fill_view(c, c->k, c->cap, kbuf);
  fill_view(c, c->v, c->cap, vbuf);
}

}  // namespace omx
