// api.cu -- the C ABI of libtfrgpu.so (include/tfrgpu.h) and the host-side runtime above the kernels:
// schema lowering, per-task decoder/encoder handles (device + stream + reusable buffers), batch
// ownership, host copies and Arrow C Data Interface export.  No torch types, no CPU fallback: every
// compute path below launches the sm_100a kernels of frame.cuh / decode.cuh / scan.cuh / encode.cuh.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "decode.cuh"
#include "encode.cuh"
#include "encode_tile.cuh"
#include "frame.cuh"
#include "host_util.h"
#include "infer.cuh"
#include "scan.cuh"
#include "tile.cuh"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int32_t fail(int32_t code, const std::string& msg) { g_last_error = msg; return code; }
#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail(TFR_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

int32_t DevBuf::ensure(size_t bytes) {
  cudaError_t e = ensure_raw(bytes);
  if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? TFR_E_OOM : TFR_E_CUDA, std::string("device allocation failed: ") + cudaGetErrorString(e));
  return TFR_OK;
}
#define TRY(expr) do { int32_t _rc = (expr); if (_rc) return _rc; } while (0)

extern "C" int32_t tfr_abi_version(void) { return TFR_ABI_VERSION; }
extern "C" const char* tfr_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* tfr_status_string(int32_t s) {
  switch (s) {
    case TFR_OK: return "ok";
    case TFR_E_INVALID_ARG: return "invalid argument";
    case TFR_E_UNSUPPORTED_TYPE: return "unsupported data type";
    case TFR_E_BAD_RECORD_TYPE: return "unsupported recordType: recordType can be ByteArray, Example or SequenceExample";
    case TFR_E_CUDA: return "CUDA error / no usable device";
    case TFR_E_OOM: return "out of memory";
    case TFR_E_BATCH_TOO_LARGE: return "batch too large (2 GiB of framed bytes / int32 Arrow offsets)";
    case TFR_E_CRC_LENGTH: return "Length header crc32 checking failed";
    case TFR_E_CRC_DATA: return "Data crc32 checking failed";
    case TFR_E_TRUNCATED: return "End of file reached before reading fully";
    case TFR_E_RECORD_TOO_LARGE: return "Record size exceeds max value of int32";
    case TFR_E_MALFORMED_PROTO: return "Protocol message was malformed";
    case TFR_E_KIND_MISMATCH: return "Feature must be of the list kind the column type requires";
    case TFR_E_EMPTY_SCALAR: return "head of empty list";
    case TFR_E_NULL_IN_NONNULL: return "field does not allow null values";
    case TFR_E_BAD_NESTING: return "Cannot convert Feature/FeatureList to this array nesting";
    default: return "unknown status";
  }
}

// ---------------------------------------------------------------------------------------------
// per-device context: CRC tables
// ---------------------------------------------------------------------------------------------
struct DeviceCtx {
  std::once_flag once;
  CrcTables* d_tabs = nullptr;
  cudaError_t err = cudaSuccess;
  int sm_count = 148;
  int max_smem_optin = 48 * 1024;
};
static DeviceCtx g_ctx[64];

static void build_crc_tables(CrcTables& t) {
  const uint32_t POLY = 0x82F63B78u;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (POLY & (0u - (c & 1u)));
    t.t0[i] = c;
  }
  auto mulmod = [&](uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (int i = 31; i >= 0; --i) {
      if ((a >> i) & 1u) p ^= b;
      b = (b >> 1) ^ (POLY & (0u - (b & 1u)));
    }
    return p;
  };
  auto xpow_bytes = [&](uint32_t nbytes) {     // x^(8*nbytes) mod P
    uint32_t x8 = 0x80000000u;
    for (int i = 0; i < 8; ++i) x8 = (x8 >> 1) ^ (POLY & (0u - (x8 & 1u)));
    uint32_t r = 0x80000000u, base = x8;
    while (nbytes) { if (nbytes & 1) r = mulmod(r, base); base = mulmod(base, base); nbytes >>= 1; }
    return r;
  };
  uint32_t x128 = xpow_bytes(128);
  for (int s = 0; s < 4; ++s)
    for (uint32_t b = 0; b < 256; ++b) t.k128[s][b] = mulmod(x128, b << (8 * s));
  for (uint32_t k = 0; k < 40; ++k) t.xw[k] = xpow_bytes(4 * k);
  for (uint32_t i = 0; i < 256; ++i) {
    t.s8[0][i] = t.t0[i];
    for (int k = 1; k < 8; ++k) t.s8[k][i] = (t.s8[k - 1][i] >> 8) ^ t.t0[t.s8[k - 1][i] & 0xff];
  }
  for (uint32_t m = 0; m < 512; ++m) t.xp16[m] = xpow_bytes(16 * m);
  memset(t.g5, 0, sizeof t.g5);
  for (uint32_t k = 0; k < 13; ++k)
    for (uint32_t v = 0; v < 32; ++v) {
      uint32_t r = 0;
      for (uint32_t j = 0; j < 5; ++j) {
        const uint32_t bit = 5 * k + j;
        if (((v >> j) & 1u) && bit < 64) r ^= t.s8[7 - bit / 8][1u << (bit % 8)];
      }
      t.g5[k * 32 + v] = r;
    }
  for (uint32_t v = 0; v < 32; ++v) t.g5[416 + v] = t.t0[v];
  for (uint32_t w = 0; w < 8; ++w) t.g5[448 + w] = t.t0[w << 5];
}

static int32_t get_ctx(int device, DeviceCtx** out) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(TFR_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev || device >= 64) return fail(TFR_E_INVALID_ARG, "bad device index");
  DeviceCtx& c = g_ctx[device];
  std::call_once(c.once, [&] {
    c.err = cudaSetDevice(device);
    if (c.err != cudaSuccess) return;
    cudaDeviceProp prop;
    c.err = cudaGetDeviceProperties(&prop, device);
    if (c.err != cudaSuccess) return;
    c.sm_count = prop.multiProcessorCount;
    c.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    {   // keep stream-ordered allocations cached in the pool instead of returning them to the OS at every sync
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
      }
    }
    CrcTables* h = new CrcTables;
    build_crc_tables(*h);
    c.err = cudaMalloc(&c.d_tabs, sizeof(CrcTables));
    if (c.err == cudaSuccess) c.err = cudaMemcpy(c.d_tabs, h, sizeof(CrcTables), cudaMemcpyHostToDevice);
    delete h;
  });
  if (c.err != cudaSuccess) return fail(TFR_E_CUDA, std::string("device init: ") + cudaGetErrorString(c.err));
  *out = &c;
  return TFR_OK;
}

// ---------------------------------------------------------------------------------------------
// schema
// ---------------------------------------------------------------------------------------------
struct tfr_schema {
  int32_t record_type = 0;
  std::vector<DevField> fields;      // device layout, filled on the host
  std::vector<uint8_t> names;
  std::vector<int32_t> ht;
  std::vector<int32_t> var_field;    // var slot -> field
  std::vector<int32_t> fix_field;    // fix slot -> field
  int32_t n_fix = 0, n_var = 0, n_cnt = 0;
};

static uint32_t fnv1a(const uint8_t* p, uint32_t n) { uint32_t h = 2166136261u; for (uint32_t i = 0; i < n; ++i) h = (h ^ p[i]) * 16777619u; return h; }

extern "C" int32_t tfr_schema_create(const tfr_field* fields, int32_t n_fields, int32_t record_type, tfr_schema** out) {
  if (!out || n_fields < 0 || (n_fields > 0 && !fields)) return fail(TFR_E_INVALID_ARG, "null argument");
  if (record_type < TFR_RT_EXAMPLE || record_type > TFR_RT_BYTE_ARRAY)
    return fail(TFR_E_BAD_RECORD_TYPE, "Unsupported recordType: recordType can be ByteArray, Example or SequenceExample");
  if (n_fields > 4096) return fail(TFR_E_INVALID_ARG, "more than 4096 fields");
  auto* s = new tfr_schema;
  s->record_type = record_type;
  if (record_type == TFR_RT_BYTE_ARRAY) {
    // the single binary column of TensorFlowInferSchema.getSchemaForByteArray (M/TensorFlowInferSchema.scala:60-64);
    // the caller's field list is ignored like deserializeByteArray ignores the schema (:17-19)
    DevField d{};
    d.name_off = 0; d.name_len = 9; s->names.assign((const uint8_t*)"byteArray", (const uint8_t*)"byteArray" + 9);
    d.hash = fnv1a(s->names.data(), 9);
    d.elem_type = TFR_T_BINARY; d.depth = 0; d.nullable = 1; d.kind = K_BYTES; d.n_levels = 1; d.dup_next = -1;
    d.fix_slot = -1; d.var_slot = 0; d.cnt_slot = 0; d.width = 1;
    s->fields.push_back(d);
    s->var_field.push_back(0);
    s->n_var = 1; s->n_cnt = 1;
    s->ht.assign(2, -1);
    s->ht[d.hash & 1] = 0;
    *out = s;
    return TFR_OK;
  }
  for (int32_t i = 0; i < n_fields; ++i) {
    const tfr_field& f = fields[i];
    if (f.name_len < 0 || (f.name_len > 0 && !f.name)) { delete s; return fail(TFR_E_INVALID_ARG, "bad field name"); }
    std::string nm(f.name ? f.name : "", (size_t)f.name_len);
    // newFeatureWriter / newFeatureConverter: anything but these types throws (M/TFRecordDeserializer.scala:119-123,
    // M/TFRecordSerializer.scala:147,151); ArrayType(NullType) falls into the same default branch
    bool ok_type = f.elem_type >= TFR_T_NULL && f.elem_type <= TFR_T_BINARY && f.depth >= 0 && f.depth <= 2 &&
                   !(f.elem_type == TFR_T_NULL && f.depth > 0);
    if (!ok_type) { delete s; return fail(TFR_E_UNSUPPORTED_TYPE, "field '" + nm + "': data type is not supported"); }
    DevField d{};
    d.name_off = (uint32_t)s->names.size();
    d.name_len = (uint32_t)f.name_len;
    s->names.insert(s->names.end(), (const uint8_t*)f.name, (const uint8_t*)f.name + f.name_len);
    d.hash = fnv1a((const uint8_t*)f.name, d.name_len);
    d.elem_type = (int8_t)f.elem_type; d.depth = (int8_t)f.depth; d.nullable = f.nullable ? 1 : 0;
    d.kind = (int8_t)required_kind(f.elem_type);
    bool varlen = f.elem_type == TFR_T_STRING || f.elem_type == TFR_T_BINARY;
    d.n_levels = (int16_t)(f.depth + (varlen ? 1 : 0));
    d.dup_next = -1;
    d.width = type_width(f.elem_type);
    d.fix_slot = -1; d.var_slot = -1; d.cnt_slot = -1;
    if (f.elem_type == TFR_T_NULL) { /* no storage beyond validity */ }
    else if (d.n_levels == 0) { d.fix_slot = s->n_fix++; s->fix_field.push_back(i); }
    else { d.var_slot = s->n_var++; d.cnt_slot = s->n_cnt; s->n_cnt += d.n_levels; s->var_field.push_back(i); }
    s->fields.push_back(d);
  }
  // Spark refuses duplicate column names for file sources before the reader is built
  // (SchemaUtils.checkColumnNameDuplication), so they never reach TFRecordDeserializer
  size_t hsz = 2; while (hsz < 2 * (size_t)n_fields + 2) hsz <<= 1;
  s->ht.assign(hsz, -1);
  for (int32_t i = 0; i < n_fields; ++i) {
    const DevField& d = s->fields[i];
    size_t slot = d.hash & (hsz - 1);
    while (s->ht[slot] >= 0) {
      const DevField& o = s->fields[s->ht[slot]];
      if (o.hash == d.hash && o.name_len == d.name_len && memcmp(&s->names[o.name_off], &s->names[d.name_off], d.name_len) == 0) {
        delete s; return fail(TFR_E_INVALID_ARG, "Found duplicate column(s) in the data schema");
      }
      slot = (slot + 1) & (hsz - 1);
    }
    s->ht[slot] = i;
  }
  *out = s;
  return TFR_OK;
}
extern "C" void tfr_schema_destroy(tfr_schema* s) { delete s; }
extern "C" int32_t tfr_schema_num_fields(const tfr_schema* s) { return s ? (int32_t)s->fields.size() : 0; }

// device copy of a schema
struct DevSchemaBuf {
  DevField* d_fields = nullptr; uint8_t* d_names = nullptr; int32_t* d_ht = nullptr; int32_t* d_var_field = nullptr;
  FieldTemplate* d_templates = nullptr;
  uint8_t* d_tile_consts = nullptr; uint32_t tile_consts_bytes = 0;     // tile.cuh: per-schema constants in the shared-memory layout
  DevSchema view{};
  int32_t upload(const tfr_schema& s) {
    size_t nf = s.fields.size();
    CUDA_TRY(cudaMalloc(&d_fields, std::max<size_t>(1, nf) * sizeof(DevField)));
    CUDA_TRY(cudaMalloc(&d_names, std::max<size_t>(1, s.names.size())));
    CUDA_TRY(cudaMalloc(&d_ht, s.ht.size() * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&d_var_field, std::max<size_t>(1, s.var_field.size()) * sizeof(int32_t)));
    if (nf) CUDA_TRY(cudaMemcpy(d_fields, s.fields.data(), nf * sizeof(DevField), cudaMemcpyHostToDevice));
    if (!s.names.empty()) CUDA_TRY(cudaMemcpy(d_names, s.names.data(), s.names.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_ht, s.ht.data(), s.ht.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (!s.var_field.empty()) CUDA_TRY(cudaMemcpy(d_var_field, s.var_field.data(), s.var_field.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    {
      // canonical entry prefix of every field: 0A ? 0A klen key 12 ? kindtag ?   (? = length bytes, masked out)
      std::vector<FieldTemplate> tp(std::max<size_t>(1, nf));
      for (size_t f = 0; f < nf; ++f) {
        FieldTemplate& t = tp[f];
        memset(&t, 0, sizeof t);
        const DevField& fd = s.fields[f];
        t.kind = (uint32_t)fd.kind;
        const uint32_t klen = fd.name_len, total = klen + 8;
        if (fd.kind == K_NONE || klen >= 0x80 || total > TILE_TPL_WORDS * 4) continue;      // no template: generic parse
        uint8_t bytes[TILE_TPL_WORDS * 4] = {0}, mask[TILE_TPL_WORDS * 4] = {0};
        auto put = [&](uint32_t i, uint8_t b, bool fixed) { bytes[i] = b; mask[i] = fixed ? 0xFF : 0x00; };
        put(0, 0x0A, true); put(1, 0, false); put(2, 0x0A, true); put(3, (uint8_t)klen, true);
        for (uint32_t i = 0; i < klen; ++i) put(4 + i, s.names[fd.name_off + i], true);
        put(4 + klen, 0x12, true); put(5 + klen, 0, false);
        put(6 + klen, fd.kind == K_BYTES ? 0x0A : fd.kind == K_FLOAT ? 0x12 : 0x1A, true); put(7 + klen, 0, false);
        t.n_words = (uint16_t)((total + 3) / 4); t.klen = (uint16_t)klen;
        memcpy(t.words, bytes, sizeof bytes); memcpy(t.mask, mask, sizeof mask);
      }
      CUDA_TRY(cudaMalloc(&d_templates, tp.size() * sizeof(FieldTemplate)));
      CUDA_TRY(cudaMemcpy(d_templates, tp.data(), tp.size() * sizeof(FieldTemplate), cudaMemcpyHostToDevice));
    }
    view.n_fields = (int32_t)nf; view.record_type = s.record_type; view.ht_mask = (int32_t)s.ht.size() - 1;
    view.n_fix = s.n_fix; view.n_var = s.n_var; view.n_cnt = s.n_cnt;
    view.fields = d_fields; view.names = d_names; view.ht = d_ht;
    return TFR_OK;
  }
  // CRC tables | zeroed seen words | DevField[nf] | FieldTemplate[nf] | names, each section 16-byte aligned (tile_const_bytes)
  int32_t build_tile_consts(const tfr_schema& s, const CrcTables* d_tabs) {
    const uint32_t nf = (uint32_t)s.fields.size(), nb = (uint32_t)s.names.size();
    tile_consts_bytes = tile_const_bytes(nf, nb);
    CUDA_TRY(cudaMalloc(&d_tile_consts, tile_consts_bytes));
    CUDA_TRY(cudaMemset(d_tile_consts, 0, tile_consts_bytes));
    uint8_t* q = d_tile_consts;
    CUDA_TRY(cudaMemcpy(q, d_tabs->g5, TILE_CRC_BYTES, cudaMemcpyDeviceToDevice));   // g5 then xp16, contiguous in CrcTables
    q += TILE_CRC_BYTES + TILE_SEEN_BYTES;
    if (nf) CUDA_TRY(cudaMemcpy(q, d_fields, nf * sizeof(DevField), cudaMemcpyDeviceToDevice));
    q += (nf * sizeof(DevField) + 15) & ~(size_t)15;
    if (nf) CUDA_TRY(cudaMemcpy(q, d_templates, nf * sizeof(FieldTemplate), cudaMemcpyDeviceToDevice));
    q += (nf * sizeof(FieldTemplate) + 15) & ~(size_t)15;
    if (nb) CUDA_TRY(cudaMemcpy(q, d_names, nb, cudaMemcpyDeviceToDevice));
    return TFR_OK;
  }
  void free_all() { cudaFree(d_fields); cudaFree(d_names); cudaFree(d_ht); cudaFree(d_var_field); cudaFree(d_templates); cudaFree(d_tile_consts); }
};

// ---------------------------------------------------------------------------------------------
// decoder
// ---------------------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct HostStats {      // pinned, written by D2H copies
  FrameResult frame;
  DecodeSummary summary;
  uint32_t overflow;
  uint32_t pad[3];
};

struct tfr_decoder {
  std::atomic<int> refs{1};            // the creator + every live batch (exported Arrow arrays keep batches alive)
  tfr_schema schema;
  int device = 0;
  uint32_t flags = 0;
  DeviceCtx* ctx = nullptr;
  cudaStream_t stream = nullptr;
  DevSchemaBuf dsch;
  // reusable device scratch
  DevBuf in, chunks, chunk_base, chunk_cnt, k1_tsum, k1_first, k1_stage, rec_off, status, valid8, cnt, src, cflag, tsum, scan_scratch, ptr_tables, small;
  // pinned host
  void* staging = nullptr; size_t staging_cap = 0;
  HostStats* h_stats = nullptr;
  int64_t* h_totals = nullptr;          // [n_cnt] + null counts [nf]
  void** h_ptr_tables = nullptr;        // pinned mirror of the device pointer tables
  size_t h_ptr_cap = 0;
  PinnedPool host_pool;
  DevPool dev_pool;                     // output blocks of the batches
  std::vector<cudaEvent_t> done_pool;
  // fast path (tile.cuh)
  bool fast_ok = false;
  size_t tile_smem_set = 0;
  int spec_state = 0;                   // 0 learning, 1 speculating on uniform shapes, -1 disabled
  double mean_rec_bytes = 0.0;          // framed bytes per record of the previous batch (frame index chunk size)
  std::vector<int32_t> spec_len;
  DevBuf uniform_dev;
  int32_t* h_uniform = nullptr;         // pinned
  void* h_k1 = nullptr;                 // pinned, 16 bytes
  // profiling (bench.py): CUDA events around the stages
  bool profiling = false;
  struct Span { int stage; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> ev_pool;
  double prof_ms[TFR_PROFILE_STAGES] = {0};
  int64_t launches = 0, pass1_launches = 0;
  cudaEvent_t ev_get() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void span_begin(int stage) { if (!profiling) return; Span s{stage, ev_get(), nullptr}; cudaEventRecord(s.a, stream); spans.push_back(s); }
  void span_end(int nlaunch) { launches += nlaunch; if (!profiling) return; Span& s = spans.back(); s.b = ev_get(); cudaEventRecord(s.b, stream); }
  void spans_resolve() {
    for (auto& s : spans) {
      if (!s.b) { ev_pool.push_back(s.a); continue; }
      float ms = 0; if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) prof_ms[s.stage] += ms;
      ev_pool.push_back(s.a); ev_pool.push_back(s.b);
    }
    spans.clear();
  }
};

struct Segment { void* dev; size_t bytes; size_t host_off; };

struct tfr_batch {
  tfr_decoder* dec = nullptr;
  std::atomic<int> refs{1};
  tfr_batch_info info{};
  std::vector<tfr_column> cols;          // device view
  std::vector<tfr_column> host_cols;     // host view (after to_host)
  void* dev_fixed = nullptr; size_t dev_fixed_bytes = 0;
  void* dev_var = nullptr; size_t dev_var_bytes = 0;
  unsigned long long* d_null_counts = nullptr;   // inside dev_fixed
  std::vector<unsigned long long> h_null_counts_tmp;
  void* host_copy = nullptr; size_t host_copy_bytes = 0;
  bool null_counts_ready = false;
  cudaEvent_t done = nullptr;
};

extern "C" int32_t tfr_decoder_create(const tfr_schema* schema, int32_t device, uint32_t flags, tfr_decoder** out) {
  if (!schema || !out) return fail(TFR_E_INVALID_ARG, "null argument");
  DeviceCtx* ctx = nullptr;
  int32_t rc = get_ctx(device, &ctx);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(device));
  auto* d = new tfr_decoder;
  d->schema = *schema; d->device = device; d->flags = flags; d->ctx = ctx;
  CUDA_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  rc = d->dsch.upload(d->schema);
  if (!rc) rc = d->dsch.build_tile_consts(d->schema, ctx->d_tabs);
  if (rc) { delete d; return rc; }
  TRY(d->small.ensure(4096 + (size_t)d->schema.n_cnt * 16));
  {
    // fast path eligibility (tile.cuh): Example records, scalars and 1-D arrays, at most 128 fields
    bool ok = d->schema.record_type != TFR_RT_BYTE_ARRAY && d->schema.fields.size() <= 128 && !getenv("TFR_DISABLE_FAST");
    for (const DevField& f : d->schema.fields) if (f.depth > 1 && d->schema.record_type == TFR_RT_EXAMPLE) ok = false;
    d->fast_ok = ok;
    if (getenv("TFR_DISABLE_SPECULATION")) d->spec_state = -1;
    d->spec_len.assign(std::max(1, d->schema.n_var), -1);
    CUDA_TRY(cudaHostAlloc((void**)&d->h_uniform, std::max<size_t>(1, d->schema.n_var) * 4, cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&d->h_k1, 64, cudaHostAllocDefault));
  }
  CUDA_TRY(cudaHostAlloc((void**)&d->h_stats, sizeof(HostStats), cudaHostAllocDefault));
  size_t nt = (size_t)d->schema.n_cnt * 2 + d->schema.fields.size() + 8;
  CUDA_TRY(cudaHostAlloc((void**)&d->h_totals, nt * sizeof(int64_t), cudaHostAllocDefault));
  *out = d;
  return TFR_OK;
}
static void decoder_unref(tfr_decoder* d);
extern "C" void tfr_decoder_destroy(tfr_decoder* d) { if (d) decoder_unref(d); }
static void decoder_unref(tfr_decoder* d) {
  if (d->refs.fetch_sub(1) != 1) return;
  cudaSetDevice(d->device);
  cudaStreamSynchronize(d->stream);
  for (DevBuf* b : {&d->in, &d->chunks, &d->chunk_base, &d->chunk_cnt, &d->k1_tsum, &d->k1_first, &d->k1_stage, &d->rec_off, &d->status, &d->valid8, &d->cnt, &d->src, &d->cflag, &d->tsum,
                    &d->scan_scratch, &d->ptr_tables, &d->small, &d->uniform_dev})
    b->release();
  d->dsch.free_all();
  if (d->staging) cudaFreeHost(d->staging);
  cudaFreeHost(d->h_stats); cudaFreeHost(d->h_totals);
  if (d->h_uniform) cudaFreeHost(d->h_uniform);
  if (d->h_k1) cudaFreeHost(d->h_k1);
  if (d->h_ptr_tables) cudaFreeHost(d->h_ptr_tables);
  d->host_pool.release_all();
  d->dev_pool.release_all();
  for (cudaEvent_t e : d->done_pool) cudaEventDestroy(e);
  d->spans_resolve();
  for (cudaEvent_t e : d->ev_pool) cudaEventDestroy(e);
  cudaStreamDestroy(d->stream);
  delete d;
}
extern "C" int32_t tfr_decoder_staging(tfr_decoder* d, size_t min_bytes, void** host_ptr, size_t* capacity) {
  if (!d || !host_ptr) return fail(TFR_E_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(d->device));
  if (d->staging_cap < min_bytes) {
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    if (d->staging) cudaFreeHost(d->staging);
    d->staging = nullptr; d->staging_cap = 0;
    size_t cap = align_up(std::max<size_t>(min_bytes, 1 << 20), 1 << 20);
    CUDA_TRY(cudaHostAlloc(&d->staging, cap, cudaHostAllocDefault));
    d->staging_cap = cap;
  }
  *host_ptr = d->staging;
  if (capacity) *capacity = d->staging_cap;
  return TFR_OK;
}
extern "C" int32_t tfr_decoder_set_profiling(tfr_decoder* d, int32_t enable) {
  if (!d) return TFR_E_INVALID_ARG;
  cudaSetDevice(d->device);
  cudaStreamSynchronize(d->stream);
  d->spans_resolve();
  d->profiling = enable != 0;
  for (double& m : d->prof_ms) m = 0;
  d->launches = 0; d->pass1_launches = 0;
  return TFR_OK;
}
extern "C" int32_t tfr_decoder_get_profile(tfr_decoder* d, double* ms, int64_t* launches, int64_t* pass1) {
  if (!d || !ms) return TFR_E_INVALID_ARG;
  CUDA_TRY(cudaSetDevice(d->device));
  CUDA_TRY(cudaStreamSynchronize(d->stream));
  d->spans_resolve();
  for (int i = 0; i < TFR_PROFILE_STAGES; ++i) ms[i] = d->prof_ms[i];
  if (launches) *launches = d->launches;
  if (pass1) *pass1 = d->pass1_launches;
  return TFR_OK;
}
extern "C" int32_t tfr_decoder_stream(tfr_decoder* d, void** s) { if (!d || !s) return TFR_E_INVALID_ARG; *s = d->stream; return TFR_OK; }

static int32_t frame_stop_to_error(uint32_t stop, bool is_final) {
  switch (stop) {
    case FS_BAD_CRC: return TFR_E_CRC_LENGTH;
    case FS_TOO_LARGE: return TFR_E_RECORD_TOO_LARGE;
    case FS_PART_HDR: case FS_PART_REC: return is_final ? TFR_E_TRUNCATED : TFR_OK;
    default: return TFR_OK;       // FS_EOF, FS_STRAY (EOFException on the length bytes is a clean EOF)
  }
}

static uint32_t pick_chunk_bytes(size_t nbytes, int sm_count, double mean_rec_bytes = 0.0) {
  // enough chunks to give every resident warp work, large enough to amortise the candidate search
  // the frame index costs one candidate search per chunk (the dominant part) plus one dependent DRAM hop per
  // record: aim for about one chunk per resident warp
  size_t want = (size_t)sm_count * 64;
  size_t c = 4096;
  while (c < 262144 && nbytes / c > want) c <<= 1;
  // the chain walk costs one dependent DRAM hop per record, the candidate search about one record length per chunk: with
  // the record size of the previous batch known, keep the chains at about 48 records (small records -> more, smaller chunks)
  if (mean_rec_bytes > 0) {
    size_t by_rec = 4096;
    while (by_rec < 262144 && (double)by_rec < 48.0 * mean_rec_bytes) by_rec <<= 1;
    while (by_rec < 262144 && nbytes / by_rec > 32768) by_rec <<= 1;
    c = std::min(c, by_rec);
  }
  return (uint32_t)c;
}

// everything one tfr_decode call needs across its stages
struct DecodeCtx {
  tfr_decoder* d; tfr_batch* b; cudaStream_t st;
  const uint8_t* d_data; size_t nbytes;
  uint32_t n = 0, n_chunks = 0, chunk_bytes = 0, verify = 0, nf = 0, nb_stride = 0;
  std::vector<size_t> fix_off, off0_off;
  size_t bitmaps_off = 0, nullc_off = 0;
  uint8_t* fx = nullptr;
  void **t_fix, **t_scan, **t_offs, **t_vals, **dt_fix, **dt_scan, **dt_offs, **dt_vals;
  bool rec_off_ready = false;
};

static int32_t alloc_fixed(DecodeCtx& C) {
  tfr_decoder* d = C.d; const tfr_schema& S = d->schema; const uint32_t n = C.n, nf = C.nf;
  C.nb_stride = (uint32_t)align_up(((size_t)n + 7) / 8 + 1, 64);
  C.fix_off.assign(S.n_fix, 0); C.off0_off.assign(S.n_var, 0);
  size_t fixed_bytes = 0;
  C.bitmaps_off = 0; fixed_bytes += align_up((size_t)C.nb_stride * std::max<uint32_t>(nf, 1), 256);
  C.nullc_off = fixed_bytes; fixed_bytes += align_up(sizeof(unsigned long long) * std::max<uint32_t>(nf, 1), 256);
  for (int i = 0; i < S.n_fix; ++i) { C.fix_off[i] = fixed_bytes; fixed_bytes += align_up((size_t)n * S.fields[S.fix_field[i]].width + 8, 256); }
  for (int v = 0; v < S.n_var; ++v) { C.off0_off[v] = fixed_bytes; fixed_bytes += align_up(((size_t)n + 1) * 4, 256); }
  C.b->dev_fixed = C.d->dev_pool.acquire(fixed_bytes);
  if (!C.b->dev_fixed) return fail(TFR_E_OOM, "device allocation failed (batch outputs)");
  C.b->dev_fixed_bytes = fixed_bytes;
  C.fx = (uint8_t*)C.b->dev_fixed;
  C.b->d_null_counts = (unsigned long long*)(C.fx + C.nullc_off);
  CUDA_TRY(cudaMemsetAsync(C.fx + C.nullc_off, 0, sizeof(unsigned long long) * std::max<uint32_t>(nf, 1), C.st));
  if (n == 0) for (int v = 0; v < S.n_var; ++v) CUDA_TRY(cudaMemsetAsync(C.fx + C.off0_off[v], 0, 4, C.st));
  // pointer tables
  const size_t n_ptr = (size_t)S.n_fix + (size_t)S.n_cnt + (size_t)S.n_var * 4 + 8;
  if (d->h_ptr_cap < n_ptr) {
    if (d->h_ptr_tables) cudaFreeHost(d->h_ptr_tables);
    CUDA_TRY(cudaHostAlloc((void**)&d->h_ptr_tables, n_ptr * sizeof(void*), cudaHostAllocDefault));
    d->h_ptr_cap = n_ptr;
  }
  TRY(d->ptr_tables.ensure(n_ptr * sizeof(void*)));
  void** hp = d->h_ptr_tables; void** dp = (void**)d->ptr_tables.p;
  C.t_fix = hp;                        C.dt_fix = dp;
  C.t_scan = hp + S.n_fix;             C.dt_scan = dp + S.n_fix;
  C.t_offs = C.t_scan + S.n_cnt;       C.dt_offs = C.dt_scan + S.n_cnt;
  C.t_vals = C.t_offs + S.n_var * 3;   C.dt_vals = C.dt_offs + S.n_var * 3;
  return TFR_OK;
}

// scratch sized by n + scan output pointers + fixed value pointers, uploaded to the device tables
static int32_t prepare_scratch(DecodeCtx& C) {
  tfr_decoder* d = C.d; const tfr_schema& S = d->schema; const uint32_t n = C.n, nf = C.nf;
  TRY(d->status.ensure((size_t)n * 4));
  TRY(d->valid8.ensure((size_t)n * std::max<uint32_t>(nf, 1) + 8));
  TRY(d->cnt.ensure((size_t)n * std::max<int>(S.n_cnt, 1) * 4));
  TRY(d->src.ensure((size_t)n * std::max<int>(S.n_var, 1) * 4));
  TRY(d->cflag.ensure((size_t)n * std::max<int>(S.n_var, 1)));
  size_t deep = 0;
  for (int v = 0; v < S.n_var; ++v) deep += (size_t)(S.fields[S.var_field[v]].n_levels - 1);
  TRY(d->scan_scratch.ensure(deep * align_up(((size_t)n + 1) * 4, 256) + 256));
  size_t so = 0;
  for (int v = 0; v < S.n_var; ++v) {
    const DevField& fd = S.fields[S.var_field[v]];
    C.t_scan[fd.cnt_slot] = C.fx + C.off0_off[v];
    for (int l = 1; l < fd.n_levels; ++l) { C.t_scan[fd.cnt_slot + l] = (uint8_t*)d->scan_scratch.p + so; so += align_up(((size_t)n + 1) * 4, 256); }
  }
  for (int i = 0; i < S.n_fix; ++i) C.t_fix[i] = C.fx + C.fix_off[i];
  CUDA_TRY(cudaMemcpyAsync(C.dt_fix, C.t_fix, ((size_t)S.n_fix + S.n_cnt) * sizeof(void*), cudaMemcpyHostToDevice, C.st));
  return TFR_OK;
}

static void fill_decode_args(DecodeCtx& C, DecodeArgs& A) {
  tfr_decoder* d = C.d;
  A = DecodeArgs{};
  A.data = C.d_data; A.rec_off = (const uint32_t*)d->rec_off.p; A.n = C.n; A.nbytes = (uint32_t)C.nbytes; A.verify = C.verify;
  A.sch = d->dsch.view; A.tabs = d->ctx->d_tabs;
  A.status = (uint32_t*)d->status.p; A.valid8 = (uint8_t*)d->valid8.p; A.fix_values = (void* const*)C.dt_fix;
  A.cnt = (uint32_t*)d->cnt.p; A.src = (uint32_t*)d->src.p; A.cflag = (uint8_t*)d->cflag.p;
  A.scan = (const int32_t* const*)C.dt_scan; A.offs = (int32_t* const*)C.dt_offs; A.var_values = (void* const*)C.dt_vals;
  A.var_field = d->dsch.d_var_field;
}

static int32_t ensure_rec_off(DecodeCtx& C) {
  if (C.rec_off_ready) return TFR_OK;
  tfr_decoder* d = C.d;
  TRY(d->rec_off.ensure(((size_t)C.n + 1) * 4));
  d->span_begin(0);
  frame_emit_kernel<<<(C.n_chunks * FRAME_EMIT_LANES + 255) / 256, 256, 0, C.st>>>(C.d_data, (const ChunkInfo*)d->chunks.p, (const uint32_t*)d->chunk_base.p, C.n_chunks,
                                                                                  (const uint32_t*)d->k1_stage.p, (const FrameResult*)d->small.p, (uint32_t*)d->rec_off.p);
  d->span_end(1);
  C.rec_off_ready = true;
  return TFR_OK;
}

// device scratch inside d->small: [0] FrameResult, [256] DecodeSummary, [512] overflow, [516] tile flags, [1024] totals, then first counts
static DecodeSummary* dsum_ptr(tfr_decoder* d) { return (DecodeSummary*)((uint8_t*)d->small.p + 256); }
static uint32_t* dovf_ptr(tfr_decoder* d) { return (uint32_t*)((uint8_t*)d->small.p + 512); }
static uint32_t* dflags_ptr(tfr_decoder* d) { return (uint32_t*)((uint8_t*)d->small.p + 516); }
static int64_t* dtot_ptr(tfr_decoder* d) { return (int64_t*)((uint8_t*)d->small.p + 1024); }

// scans + summary + the one D2H that carries totals / first error / flags; leaves results in h_stats / h_totals
static int32_t scans_and_sync(DecodeCtx& C, DecodeArgs& A, bool with_status) {
  tfr_decoder* d = C.d; const tfr_schema& S = d->schema; const uint32_t n = C.n; cudaStream_t st = C.st;
  DecodeSummary init{}; init.first_err_row = 0xffffffffu; init.n_eff = n;
  d->h_stats->summary = init; d->h_stats->overflow = 0;
  CUDA_TRY(cudaMemcpyAsync(dsum_ptr(d), &d->h_stats->summary, sizeof(DecodeSummary), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(dovf_ptr(d), 0, 4, st));
  d->span_begin(2);
  int nl = 0;
  if (with_status) { first_error_kernel<<<std::min<uint32_t>((n + 255) / 256, 1024), 256, 0, st>>>(A.status, n, dsum_ptr(d)); ++nl; }
  if (S.n_cnt > 0) {
    uint32_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    TRY(d->tsum.ensure(((size_t)S.n_cnt * n_tiles + S.n_cnt) * 8 + 64));
    uint64_t* tsum = (uint64_t*)d->tsum.p;
    uint64_t* traw = tsum + (size_t)S.n_cnt * n_tiles;
    scan_tile_sums_kernel<<<dim3(n_tiles, S.n_cnt), SCAN_THREADS, 0, st>>>(A.cnt, n, n_tiles, tsum);
    scan_tile_bases_kernel<<<S.n_cnt, 1024, 0, st>>>(tsum, n_tiles, traw, dovf_ptr(d));
    scan_apply_kernel<<<dim3(n_tiles, S.n_cnt), SCAN_THREADS, 0, st>>>(A.cnt, n, n_tiles, tsum, traw, (int32_t* const*)C.dt_scan);
    nl += 3;
  }
  summary_kernel<<<1, 256, 0, st>>>(with_status ? A.status : nullptr, with_status ? A.rec_off : nullptr, n, dsum_ptr(d), (const int32_t* const*)C.dt_scan, (uint32_t)S.n_cnt,
                                    dtot_ptr(d), dtot_ptr(d) + S.n_cnt);
  d->span_end(nl + 1);
  CUDA_TRY(cudaMemcpyAsync(&d->h_stats->summary, dsum_ptr(d), sizeof(DecodeSummary), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&d->h_stats->overflow, dovf_ptr(d), 8, cudaMemcpyDeviceToHost, st));   // overflow + tile flags
  if (S.n_cnt) CUDA_TRY(cudaMemcpyAsync(d->h_totals, dtot_ptr(d), (size_t)S.n_cnt * 16, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  return TFR_OK;
}

// variable-width outputs sized from totals, pass 2, validity pack, column views
static int32_t finish_var_and_views(DecodeCtx& C, DecodeArgs& A, uint32_t n_eff, const int64_t* totals, bool run_pass2, bool pack = true) {
  tfr_decoder* d = C.d; const tfr_schema& S = d->schema; tfr_batch* b = C.b; cudaStream_t st = C.st; const uint32_t n = C.n, nf = C.nf;
  std::vector<size_t> lvl_off((size_t)S.n_var * 3, 0), val_off(S.n_var, 0);
  uint8_t* vx = (uint8_t*)b->dev_var;
  if (run_pass2) {
    size_t var_bytes = 0;
    for (int v = 0; v < S.n_var; ++v) {
      const DevField& fd = S.fields[S.var_field[v]];
      for (int l = 1; l < fd.n_levels; ++l) { lvl_off[v * 3 + l] = var_bytes; var_bytes += align_up(((size_t)totals[fd.cnt_slot + l - 1] + 1) * 4, 256); }
      val_off[v] = var_bytes; var_bytes += align_up((size_t)totals[fd.cnt_slot + fd.n_levels - 1] * fd.width + 8, 256);
    }
    if (var_bytes) { b->dev_var = d->dev_pool.acquire(var_bytes); if (!b->dev_var) return fail(TFR_E_OOM, "device allocation failed (batch outputs)"); }
    b->dev_var_bytes = var_bytes;
    vx = (uint8_t*)b->dev_var;
    for (int v = 0; v < S.n_var; ++v) {
      const DevField& fd = S.fields[S.var_field[v]];
      C.t_offs[v * 3 + 0] = C.fx + C.off0_off[v];
      for (int l = 1; l < 3; ++l) C.t_offs[v * 3 + l] = l < fd.n_levels ? vx + lvl_off[v * 3 + l] : nullptr;
      for (int l = 1; l < fd.n_levels; ++l) CUDA_TRY(cudaMemsetAsync(vx + lvl_off[v * 3 + l], 0, 4, st));
      C.t_vals[v] = vx + val_off[v];
    }
    CUDA_TRY(cudaMemcpyAsync(C.dt_offs, C.t_offs, ((size_t)S.n_var * 4) * sizeof(void*), cudaMemcpyHostToDevice, st));
    A.n_eff = n_eff;
    if (n_eff > 0 && S.n_var > 0) {
      const uint32_t warps = 8;
      unsigned long long cells = S.record_type == TFR_RT_BYTE_ARRAY ? (unsigned long long)n_eff * 32 : (unsigned long long)n_eff * S.n_var;
      uint32_t g2 = (uint32_t)std::min<unsigned long long>((cells + 255) / 256, (unsigned long long)d->ctx->sm_count * 16);
      g2 = std::max<uint32_t>(g2, 1);
      d->span_begin(3);
      int nl2 = 1;
      if (A.flist_warp) {      // canonical FeatureList cells with fixed-width elements: one warp per cell
        const unsigned long long wcells = (unsigned long long)n_eff * S.n_var;
        const uint32_t gw = (uint32_t)std::min<unsigned long long>((wcells + warps - 1) / warps, (unsigned long long)d->ctx->sm_count * 32);
        decode_pass2_flist_kernel<<<std::max<uint32_t>(gw, 1), warps * 32, 0, st>>>(A);
        ++nl2;
      }
      if (S.record_type != TFR_RT_BYTE_ARRAY && !getenv("TFR_DISABLE_CANON_LEAN")) {   // the ordinary cells: small kernel, many resident threads
        A.canon_lean = 1;
        const unsigned long long wgroups = (unsigned long long)((n_eff + 31) / 32) * S.n_var;      // one warp per (32 rows, column)
        const uint32_t gl = (uint32_t)std::min<unsigned long long>((wgroups + 7) / 8, (unsigned long long)d->ctx->sm_count * 48);
        decode_pass2_canon_kernel<<<std::max<uint32_t>(gl, 1), 256, 0, st>>>(A);
        ++nl2;
      }
      decode_pass2_kernel<<<g2, warps * 32, 0, st>>>(A);
      d->span_end(nl2);
    }
  } else {
    // uniform mode: the var block was allocated up front, one level per column
    size_t off = 0;
    for (int v = 0; v < S.n_var; ++v) { val_off[v] = off; off += align_up((size_t)totals[S.fields[S.var_field[v]].cnt_slot] * S.fields[S.var_field[v]].width + 8, 256); }
  }
  if (pack && n_eff > 0 && nf > 0) {
    uint32_t nbytes_bm = (n_eff + 7) / 8;
    dim3 g((nbytes_bm + 255) / 256, nf);
    g.x = std::min<uint32_t>(g.x, 4096);
    d->span_begin(4);
    pack_validity_kernel<<<g, 256, 0, st>>>((const uint8_t*)d->valid8.p, n, n_eff, nf, C.nb_stride, C.fx + C.bitmaps_off, b->d_null_counts);
    d->span_end(1);
  }
  b->cols.resize(nf);
  int64_t out_bytes = 0;
  for (uint32_t f = 0; f < nf; ++f) {
    const DevField& fd = S.fields[f];
    tfr_column c{};
    c.elem_type = fd.elem_type; c.depth = fd.depth; c.n_levels = fd.n_levels; c.value_width = fd.width;
    c.n_rows = n_eff; c.validity = C.fx + C.bitmaps_off + (size_t)f * C.nb_stride;
    out_bytes += (n_eff + 7) / 8;
    if (fd.fix_slot >= 0) { c.values = C.fx + C.fix_off[fd.fix_slot]; c.n_values = n_eff; out_bytes += (int64_t)n_eff * fd.width; }
    else if (fd.var_slot >= 0) {
      int v = fd.var_slot;
      c.offsets[0] = (int32_t*)(C.fx + C.off0_off[v]); c.n_offsets[0] = (int64_t)n_eff + 1;
      for (int l = 1; l < fd.n_levels; ++l) { c.offsets[l] = (int32_t*)(vx + lvl_off[v * 3 + l]); c.n_offsets[l] = totals[fd.cnt_slot + l - 1] + 1; }
      c.values = vx + val_off[v]; c.n_values = totals[fd.cnt_slot + fd.n_levels - 1];
      for (int l = 0; l < fd.n_levels; ++l) out_bytes += c.n_offsets[l] * 4;
      out_bytes += c.n_values * fd.width;
    }
    b->cols[f] = c;
  }
  b->info.out_bytes = out_bytes;
  return TFR_OK;
}

static int32_t decode_impl(tfr_decoder* d, const void* data, size_t nbytes, int32_t data_on_device, int32_t is_final, tfr_batch** out,
                           size_t* consumed, bool allow_fast);
extern "C" int32_t tfr_decode(tfr_decoder* d, const void* data, size_t nbytes, int32_t data_on_device, int32_t is_final, tfr_batch** out,
                              size_t* consumed) {
  return decode_impl(d, data, nbytes, data_on_device, is_final, out, consumed, true);
}
static int32_t decode_impl(tfr_decoder* d, const void* data, size_t nbytes, int32_t data_on_device, int32_t is_final, tfr_batch** out,
                           size_t* consumed, bool allow_fast) {
  if (!d || !out || (nbytes && !data)) return fail(TFR_E_INVALID_ARG, "null argument");
  if (nbytes >= (1ull << 31)) return fail(TFR_E_BATCH_TOO_LARGE, "tfr_decode: a batch must be smaller than 2 GiB; split the file at record boundaries");
  CUDA_TRY(cudaSetDevice(d->device));
  cudaStream_t st = d->stream;
  const tfr_schema& S = d->schema;
  auto* b = new tfr_batch;
  b->dec = d;
  d->refs.fetch_add(1);
  b->info.error_row = -1; b->info.error_field = -1;
  std::unique_ptr<tfr_batch, void (*)(tfr_batch*)> guard(b, [](tfr_batch* x) { tfr_batch_release(x); });
  DecodeCtx C;
  C.d = d; C.b = b; C.st = st; C.nbytes = nbytes; C.nf = (uint32_t)S.fields.size();
  C.verify = (d->flags & TFR_F_VERIFY_CRC) ? 1u : 0u;
  const uint32_t nf = C.nf;

  // ---- input ----
  C.d_data = (const uint8_t*)data;
  if (!data_on_device && nbytes) {
    TRY(d->in.ensure(align_up(nbytes + 64, 256)));
    d->span_begin(5);
    CUDA_TRY(cudaMemcpyAsync(d->in.p, data, nbytes, cudaMemcpyHostToDevice, st));
    d->span_end(0);
    C.d_data = (const uint8_t*)d->in.p;
  }
  // ---- K1: record boundaries ----
  // On the fast path the chain trusts the length fields (pure pointer chase) and the tile kernel's CRC warp
  // verifies every length CRC from shared memory; if anything is off, K1 is redone with verification.
  FrameResult fr{};
  fr.stop = FS_EOF;
  bool k1_verified = true;
  auto run_k1 = [&](bool verify_headers) -> int32_t {
    C.chunk_bytes = pick_chunk_bytes(nbytes, d->ctx->sm_count, d->mean_rec_bytes);
    C.n_chunks = (uint32_t)((nbytes + C.chunk_bytes - 1) / C.chunk_bytes);
    TRY(d->chunks.ensure((size_t)C.n_chunks * sizeof(ChunkInfo)));
    TRY(d->chunk_base.ensure(((size_t)C.n_chunks + 1) * sizeof(uint32_t)));
    FrameResult* d_fr = (FrameResult*)d->small.p;
    FrameResult init{}; init.first_bad = 0xffffffffu;
    d->h_stats->frame = init;
    CUDA_TRY(cudaMemcpyAsync(d_fr, &d->h_stats->frame, sizeof(FrameResult), cudaMemcpyHostToDevice, st));
    const uint32_t v = verify_headers ? C.verify : 0u;
    TRY(d->chunk_cnt.ensure(((size_t)C.n_chunks + 1) * 4));
    const uint32_t kt = (C.n_chunks + SCAN_TILE - 1) / SCAN_TILE;
    TRY(d->k1_tsum.ensure(((size_t)kt + 2) * 8 + 64));
    uint64_t* tsum = (uint64_t*)d->k1_tsum.p; uint64_t* traw = tsum + kt;
    uint32_t* d_stop = (uint32_t*)((uint8_t*)d->small.p + 128);                 // [0] stop chunk, [1] scratch overflow flag, [2..3] chunk_base pointer
    uint32_t stop_init[2] = {0xffffffffu, 0};
    memcpy(d->h_k1, stop_init, 8);
    void* cb = d->chunk_base.p; memcpy((uint8_t*)d->h_k1 + 8, &cb, sizeof(void*));
    CUDA_TRY(cudaMemcpyAsync(d_stop, d->h_k1, 16, cudaMemcpyHostToDevice, st));
    d->span_begin(0);
    TRY(d->k1_first.ensure(((size_t)C.n_chunks + 1) * 4));
    TRY(d->k1_stage.ensure((size_t)C.n_chunks * FRAME_STAGE * 4 + 64));
    frame_search_kernel<<<std::min<uint32_t>((C.n_chunks + 7) / 8, (uint32_t)d->ctx->sm_count * 8), 256, 0, st>>>(C.d_data, (uint32_t)nbytes, C.chunk_bytes, C.n_chunks,
                                                                                                             d->ctx->d_tabs, (uint32_t*)d->k1_first.p);
    frame_scan_kernel<<<(C.n_chunks + 127) / 128, 128, 0, st>>>(C.d_data, (uint32_t)nbytes, C.chunk_bytes, C.n_chunks, v, d->ctx->d_tabs, (const uint32_t*)d->k1_first.p,
                                                             (ChunkInfo*)d->chunks.p, (uint32_t*)d->chunk_cnt.p, (uint32_t*)d->k1_stage.p, d_fr);
    frame_check_kernel<<<(C.n_chunks + 255) / 256, 256, 0, st>>>((const ChunkInfo*)d->chunks.p, C.n_chunks, d_fr);
    frame_repair_kernel<<<1, 32, 0, st>>>(C.d_data, (uint32_t)nbytes, C.chunk_bytes, C.n_chunks, v, d->ctx->d_tabs, (ChunkInfo*)d->chunks.p,
                                          (uint32_t*)d->chunk_cnt.p, (uint32_t*)d->k1_stage.p, d_fr);
    frame_stop_kernel<<<(C.n_chunks + 255) / 256, 256, 0, st>>>((const ChunkInfo*)d->chunks.p, C.n_chunks, d_stop);
    scan_tile_sums_kernel<<<dim3(kt, 1), SCAN_THREADS, 0, st>>>((const uint32_t*)d->chunk_cnt.p, C.n_chunks, kt, tsum);
    scan_tile_bases_kernel<<<1, 1024, 0, st>>>(tsum, kt, traw, d_stop + 1);
    scan_apply_kernel<<<dim3(kt, 1), SCAN_THREADS, 0, st>>>((const uint32_t*)d->chunk_cnt.p, C.n_chunks, kt, tsum, traw, (int32_t* const*)(d_stop + 2));
    frame_finish_kernel<<<1, 32, 0, st>>>((const ChunkInfo*)d->chunks.p, C.n_chunks, (uint32_t)nbytes, (const uint32_t*)d->chunk_base.p, d_stop, d_fr);
    d->span_end(9);
    CUDA_TRY(cudaMemcpyAsync(&d->h_stats->frame, d_fr, sizeof(FrameResult), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));                       // sync #1: number of records
    CUDA_TRY(cudaGetLastError());
    fr = d->h_stats->frame;
    C.n = fr.n_records;
    if (fr.n_records > 64) d->mean_rec_bytes = (double)nbytes / fr.n_records;     // next batch's chunk size
    C.rec_off_ready = false;
    k1_verified = verify_headers || !C.verify;
    return TFR_OK;
  };
  // the tile kernel's bulk copies need 16-byte aligned global addresses: a device buffer at any other alignment is legal
  // input but takes the general path
  const bool try_fast = allow_fast && d->fast_ok && (reinterpret_cast<uintptr_t>(C.d_data) & 15u) == 0;
  if (nbytes) TRY(run_k1(!try_fast));
  if (nbytes && !k1_verified && frame_stop_to_error(fr.stop, is_final != 0) != TFR_OK) TRY(run_k1(true));   // a framing problem: get the exact verdict
  const uint32_t n = C.n;
  int32_t frame_err = frame_stop_to_error(fr.stop, is_final != 0);
  size_t used = nbytes;
  if (nbytes) {
    if (frame_err) used = fr.stop_pos;
    else if (fr.stop == FS_EOF) used = nbytes;
    else if (fr.stop == FS_STRAY) used = is_final ? nbytes : fr.stop_pos;
    else used = fr.stop_pos;                                   // partial tail of a non-final block is carried over
  }
  b->info.n_records = n; b->info.frame_repairs = (int32_t)fr.repairs;

  TRY(alloc_fixed(C));
  int64_t* totals = d->h_totals;                      // [n_cnt] totals at n_eff, then [n_cnt] first-row counts
  DecodeSummary sum{}; sum.first_err_row = 0xffffffffu; sum.n_eff = n;
  uint32_t n_eff = n;
  if (n > 0) {
    TRY(prepare_scratch(C));
    DecodeArgs A;
    fill_decode_args(C, A);
    bool done = false;
    // ================= fast path: shared-memory tiles, one record per thread =================
    const uint32_t names_bytes = (uint32_t)S.names.size();
    // one slot per record: 16-byte alignment slack + framed record + over-read slack, an odd number of 16-byte units
    size_t slot64 = align_up((size_t)fr.max_len + 16 + 15 + 32, 16);
    if (((slot64 >> 4) & 1) == 0) slot64 += 16;
    const uint32_t tile_slot = (uint32_t)std::min<size_t>(slot64, 1u << 20);
    const uint32_t tile_cap = TILE_ROWS * tile_slot;
    const size_t tile_smem = tile_smem_bytes(nf, names_bytes, tile_cap);
    if (try_fast && tile_smem <= (size_t)d->ctx->max_smem_optin) {
      const bool uniform = d->spec_state == 1;
      TRY(ensure_rec_off(C));
      fill_decode_args(C, A);
      TRY(d->uniform_dev.ensure(std::max<size_t>(1, S.n_var) * 4));
      std::vector<int32_t> ul(std::max(1, S.n_var), -1);
      std::vector<int64_t> utot(std::max(1, S.n_cnt), 0);
      if (uniform) {
        size_t var_bytes = 0;
        for (int v = 0; v < S.n_var; ++v) {
          const DevField& fd = S.fields[S.var_field[v]];
          ul[v] = d->spec_len[v];
          utot[fd.cnt_slot] = (int64_t)ul[v] * n;
          if (utot[fd.cnt_slot] > 0x7fffffffLL) return fail(TFR_E_BATCH_TOO_LARGE, "Arrow int32 offsets overflow; decode smaller blocks");
          C.t_vals[v] = (void*)var_bytes;                          // offset for now, rebased after the allocation
          var_bytes += align_up((size_t)utot[fd.cnt_slot] * fd.width + 8, 256);
        }
        if (var_bytes) { b->dev_var = d->dev_pool.acquire(var_bytes); if (!b->dev_var) return fail(TFR_E_OOM, "device allocation failed (batch outputs)"); }
        b->dev_var_bytes = var_bytes;
        for (int v = 0; v < S.n_var; ++v) {
          C.t_vals[v] = (uint8_t*)b->dev_var + (size_t)C.t_vals[v];
          C.t_offs[v * 3 + 0] = C.fx + C.off0_off[v]; C.t_offs[v * 3 + 1] = nullptr; C.t_offs[v * 3 + 2] = nullptr;
        }
        CUDA_TRY(cudaMemcpyAsync(C.dt_offs, C.t_offs, ((size_t)S.n_var * 4) * sizeof(void*), cudaMemcpyHostToDevice, st));
      }
      memcpy(d->h_uniform, ul.data(), (size_t)std::max(1, S.n_var) * 4);
      CUDA_TRY(cudaMemcpyAsync(d->uniform_dev.p, d->h_uniform, (size_t)std::max(1, S.n_var) * 4, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemsetAsync(dflags_ptr(d), 0, 4, st));
      TileArgs TA{};
      TA.data = C.d_data; TA.nbytes = (uint32_t)nbytes; TA.rec_off = (const uint32_t*)d->rec_off.p; TA.n = n; TA.tile_cap = tile_cap; TA.slot = tile_slot;
      TA.verify = C.verify; TA.names_bytes = names_bytes; TA.sch = d->dsch.view; TA.consts = d->dsch.d_tile_consts; TA.const_bytes = d->dsch.tile_consts_bytes;
      TA.bitmaps = C.fx + C.bitmaps_off; TA.nb_stride = C.nb_stride; TA.null_counts = b->d_null_counts;
      TA.fix_values = A.fix_values; TA.cnt = A.cnt; TA.src = A.src; TA.cflag = A.cflag;
      TA.uniform_len = (const int32_t*)d->uniform_dev.p; TA.var_values = (void* const*)C.dt_vals; TA.flags = dflags_ptr(d);
      const bool seq = S.record_type == TFR_RT_SEQUENCE_EXAMPLE;
      if (d->tile_smem_set < tile_smem) {
        if (seq) CUDA_TRY(cudaFuncSetAttribute(decode_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
        else CUDA_TRY(cudaFuncSetAttribute(decode_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
        d->tile_smem_set = tile_smem;
      }
      d->span_begin(1);
      if (seq) decode_tile_kernel<true><<<(n + TILE_ROWS - 1) / TILE_ROWS, TILE_THREADS, tile_smem, st>>>(TA);
      else decode_tile_kernel<false><<<(n + TILE_ROWS - 1) / TILE_ROWS, TILE_THREADS, tile_smem, st>>>(TA);
      d->span_end(1); d->pass1_launches++;
      uint32_t tflags = 0;
      if (uniform) {
        if (S.n_var) {       // Arrow offsets of uniform columns: offs[i] = i * L
          d->span_begin(2);
          uniform_offsets_kernel<<<dim3(std::min<uint32_t>((n + 256) / 256, 512), S.n_var), 256, 0, st>>>((int32_t* const*)C.dt_offs, 3, (const int32_t*)d->uniform_dev.p, (uint32_t)S.n_var, n);
          d->span_end(1);
        }
        CUDA_TRY(cudaMemcpyAsync(&d->h_stats->overflow, dovf_ptr(d), 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));                  // the only other sync of the fast path: the verdict flag
        CUDA_TRY(cudaGetLastError());
        tflags = (&d->h_stats->overflow)[1];
        if (!(tflags & TF_FALLBACK)) {
          for (int a = 0; a < S.n_cnt; ++a) totals[a] = utot[a];
          TRY(finish_var_and_views(C, A, n, totals, false, false));
          done = true;
        } else {
          if (tflags & TF_SHAPE) d->spec_state = -1;          // shapes are not uniform after all: stop speculating
          if (b->dev_var) { d->dev_pool.give_back(b->dev_var); b->dev_var = nullptr; b->dev_var_bytes = 0; }
        }
      } else {
        TRY(scans_and_sync(C, A, false));
        tflags = (&d->h_stats->overflow)[1];
        if (!(tflags & TF_FALLBACK)) {
          if (d->h_stats->overflow)
            for (int a = 0; a < S.n_cnt; ++a) if (totals[a] < 0 || totals[a] > 0x7fffffffLL) return fail(TFR_E_BATCH_TOO_LARGE, "Arrow int32 offsets overflow; decode smaller blocks");
          A.flist_warp = S.record_type == TFR_RT_SEQUENCE_EXAMPLE && !getenv("TFR_DISABLE_FLIST_WARP");   // the tile kernel validated every FeatureList as canonical
          TRY(finish_var_and_views(C, A, n, totals, true, false));
          done = true;
          // learn shapes: every variable-width column single-level and total == n * (count of row 0)
          if (d->spec_state == 0 && S.n_var > 0) {
            bool ok = true;
            for (int v = 0; v < S.n_var && ok; ++v) {
              const DevField& fd = S.fields[S.var_field[v]];
              int64_t c0 = totals[S.n_cnt + fd.cnt_slot];
              if (fd.n_levels != 1 || c0 <= 0 || totals[fd.cnt_slot] != c0 * (int64_t)n) ok = false;
              else d->spec_len[v] = (int32_t)c0;
            }
            d->spec_state = ok ? 1 : -1;
          }
        }
      }
      if (done) { n_eff = n; sum.n_eff = n; }
    }
    // ================= general path: warp per record, full protobuf semantics =================
    if (!done && !k1_verified) {
      // the fast path could not vouch for this batch and the length CRCs were never checked: start over with the
      // exact path (frame index with verification + general kernels); the input is already on the device
      const uint8_t* dev_in = C.d_data;
      guard.reset();
      return decode_impl(d, dev_in, nbytes, 1, is_final, out, consumed, false);
    }
    if (!done) {
      // (a failed fast attempt may have touched the null counters)
      CUDA_TRY(cudaMemsetAsync(b->d_null_counts, 0, sizeof(unsigned long long) * std::max<uint32_t>(nf, 1), st));
      TRY(ensure_rec_off(C));
      fill_decode_args(C, A);
      const uint32_t warps = 8;
      size_t smem1 = CRC_SMEM_WORDS * 4 + (size_t)warps * ((nf + 3) & ~3u);
      uint32_t g1 = std::min<uint32_t>((n + warps - 1) / warps, (uint32_t)d->ctx->sm_count * 16);
      d->span_begin(1);
      decode_pass1_kernel<<<g1, warps * 32, smem1, st>>>(A);
      d->span_end(1); d->pass1_launches++;
      TRY(scans_and_sync(C, A, true));
      sum = d->h_stats->summary;
      n_eff = sum.n_eff;
      if (sum.first_err_row != 0xffffffffu) used = sum.consumed;       // a failing record: the stream stops in front of it
      if (d->h_stats->overflow)
        for (int a = 0; a < S.n_cnt; ++a) if (totals[a] < 0 || totals[a] > 0x7fffffffLL) return fail(TFR_E_BATCH_TOO_LARGE, "Arrow int32 offsets overflow; decode smaller blocks");
      TRY(finish_var_and_views(C, A, n_eff, totals, true));
    }
  } else {
    // no complete record: empty columns
    b->cols.resize(nf);
    for (uint32_t f = 0; f < nf; ++f) {
      const DevField& fd = S.fields[f];
      tfr_column c{};
      c.elem_type = fd.elem_type; c.depth = fd.depth; c.n_levels = fd.n_levels; c.value_width = fd.width;
      c.validity = C.fx + C.bitmaps_off + (size_t)f * C.nb_stride;
      if (fd.var_slot >= 0) { c.offsets[0] = (int32_t*)(C.fx + C.off0_off[fd.var_slot]); c.n_offsets[0] = 1; }
      if (fd.fix_slot >= 0) c.values = C.fx + C.fix_off[fd.fix_slot];
      for (int l = 1; l < fd.n_levels; ++l) { c.offsets[l] = c.offsets[0]; c.n_offsets[l] = 1; }
      b->cols[f] = c;
    }
  }
  if (consumed) *consumed = used;
  b->info.consumed_bytes = (int64_t)used;
  // ---- first error in record order: per-row status (pass 1) or the framing stop at row n ----
  b->info.n_rows = n_eff;
  if (n > 0 && sum.first_err_row != 0xffffffffu) {
    b->info.error_code = -(int32_t)(sum.first_err_status & 0xff);
    b->info.error_row = sum.first_err_row;
    b->info.error_field = (int32_t)(sum.first_err_status >> 8) - 1;
  } else if (frame_err) {
    b->info.error_code = frame_err; b->info.error_row = n; b->info.error_field = -1;
  }
  if (!d->done_pool.empty()) { b->done = d->done_pool.back(); d->done_pool.pop_back(); }
  else CUDA_TRY(cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(b->done, st));
  CUDA_TRY(cudaGetLastError());
  guard.release();
  *out = b;
  return TFR_OK;
}

extern "C" int32_t tfr_batch_wait(tfr_batch* b) {
  if (!b) return fail(TFR_E_INVALID_ARG, "null batch");
  CUDA_TRY(cudaSetDevice(b->dec->device));
  if (b->done) CUDA_TRY(cudaEventSynchronize(b->done));
  if (!b->null_counts_ready) {
    size_t nf = b->cols.size();
    if (nf && b->info.n_rows > 0) {
      b->h_null_counts_tmp.resize(nf);
      CUDA_TRY(cudaMemcpy(b->h_null_counts_tmp.data(), b->d_null_counts, nf * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      for (size_t f = 0; f < nf; ++f) b->cols[f].null_count = (int64_t)b->h_null_counts_tmp[f];
    }
    b->null_counts_ready = true;
  }
  return TFR_OK;
}
extern "C" int32_t tfr_batch_status(tfr_batch* b, tfr_batch_info* out) {
  if (!b || !out) return fail(TFR_E_INVALID_ARG, "null argument");
  *out = b->info;
  return TFR_OK;
}
extern "C" int32_t tfr_batch_num_columns(tfr_batch* b) { return b ? (int32_t)b->cols.size() : 0; }
extern "C" int32_t tfr_batch_columns(tfr_batch* b, tfr_column* out, int32_t n) {
  if (!b || !out || n < (int32_t)b->cols.size()) return fail(TFR_E_INVALID_ARG, "bad argument");
  int32_t rc = tfr_batch_wait(b);
  if (rc) return rc;
  std::copy(b->cols.begin(), b->cols.end(), out);
  return TFR_OK;
}

// D2H of both output blocks into one pinned buffer, then host views with the same relative layout
extern "C" int32_t tfr_batch_to_host(tfr_batch* b, tfr_column* out, int32_t n) {
  if (!b || n < (int32_t)b->cols.size()) return fail(TFR_E_INVALID_ARG, "bad argument");
  tfr_decoder* d = b->dec;
  CUDA_TRY(cudaSetDevice(d->device));
  if (!b->host_copy) {
    size_t total = align_up(b->dev_fixed_bytes, 256) + align_up(b->dev_var_bytes, 256) + 256;
    void* h = d->host_pool.acquire(total);
    if (!h) return fail(TFR_E_OOM, "pinned host allocation failed");
    b->host_copy = h; b->host_copy_bytes = total;
    uint8_t* hb = (uint8_t*)h;
    if (b->dev_fixed_bytes) CUDA_TRY(cudaMemcpyAsync(hb, b->dev_fixed, b->dev_fixed_bytes, cudaMemcpyDeviceToHost, d->stream));
    if (b->dev_var_bytes) CUDA_TRY(cudaMemcpyAsync(hb + align_up(b->dev_fixed_bytes, 256), b->dev_var, b->dev_var_bytes, cudaMemcpyDeviceToHost, d->stream));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    int32_t rc = tfr_batch_wait(b);
    if (rc) return rc;
    auto xl = [&](void* p) -> void* {
      if (!p) return nullptr;
      uint8_t* q = (uint8_t*)p;
      uint8_t* f0 = (uint8_t*)b->dev_fixed; uint8_t* v0 = (uint8_t*)b->dev_var;
      if (f0 && q >= f0 && q < f0 + b->dev_fixed_bytes) return hb + (q - f0);
      if (v0 && q >= v0 && q < v0 + b->dev_var_bytes) return hb + align_up(b->dev_fixed_bytes, 256) + (q - v0);
      return nullptr;
    };
    b->host_cols = b->cols;
    for (auto& c : b->host_cols) {
      c.validity = (uint8_t*)xl(c.validity);
      for (int l = 0; l < 3; ++l) c.offsets[l] = (int32_t*)xl(c.offsets[l]);
      c.values = xl(c.values);
    }
  }
  if (out) std::copy(b->host_cols.begin(), b->host_cols.end(), out);
  return TFR_OK;
}

extern "C" void tfr_batch_release(tfr_batch* b) {
  if (!b) return;
  if (b->refs.fetch_sub(1) != 1) return;
  tfr_decoder* d = b->dec;
  cudaSetDevice(d->device);
  if (b->dev_fixed) d->dev_pool.give_back(b->dev_fixed);
  if (b->dev_var) d->dev_pool.give_back(b->dev_var);
  if (b->host_copy) d->host_pool.give_back(b->host_copy);
  if (b->done) d->done_pool.push_back(b->done);
  delete b;
  decoder_unref(d);
}

// ---------------------------------------------------------------------------------------------
// Arrow C Data Interface export (arrow/c/abi.h structs restated in host_util.h)
// ---------------------------------------------------------------------------------------------
struct ExportPriv { tfr_batch* batch; std::vector<const void*> buffers; std::vector<ArrowArray*> children; ArrowArray* child_storage = nullptr; };

static void release_array(ArrowArray* a) {
  if (!a || !a->release) return;
  auto* p = (ExportPriv*)a->private_data;
  for (int64_t i = 0; i < a->n_children; ++i) {
    if (a->children[i]->release) a->children[i]->release(a->children[i]);
    delete a->children[i];
  }
  if (p) { if (p->batch) tfr_batch_release(p->batch); delete p; }
  a->release = nullptr;
}
static void release_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  for (int64_t i = 0; i < s->n_children; ++i) {
    if (s->children[i]->release) s->children[i]->release(s->children[i]);
    delete s->children[i];
  }
  delete[] s->children;
  free((void*)s->name);
  s->release = nullptr;
}
static const char* leaf_format(int t) {
  switch (t) {
    case TFR_T_INT32: return "i"; case TFR_T_INT64: return "l"; case TFR_T_FLOAT32: return "f";
    case TFR_T_FLOAT64: case TFR_T_DECIMAL: return "g"; case TFR_T_STRING: return "u"; case TFR_T_BINARY: return "z";
    default: return "n";
  }
}
static void build_schema(ArrowSchema* s, const char* name, int elem_type, int depth) {
  memset(s, 0, sizeof *s);
  s->name = strdup(name); s->flags = 2 /*ARROW_FLAG_NULLABLE*/; s->release = release_schema;
  if (depth == 0) { s->format = leaf_format(elem_type); return; }
  s->format = "+l";
  s->n_children = 1; s->children = new ArrowSchema*[1];
  s->children[0] = new ArrowSchema;
  build_schema(s->children[0], "item", elem_type, depth - 1);
}
// level: which offsets level this list node uses; leaves use the last level for utf8/binary
static void build_array(ArrowArray* a, const tfr_column& c, int level, int64_t length, tfr_batch* owner) {
  memset(a, 0, sizeof *a);
  auto* p = new ExportPriv;
  p->batch = owner;
  if (owner) owner->refs.fetch_add(1);
  a->private_data = p; a->release = release_array; a->length = length; a->offset = 0;
  const bool varlen = c.elem_type == TFR_T_STRING || c.elem_type == TFR_T_BINARY;
  const void* validity = level == 0 ? c.validity : nullptr;
  a->null_count = level == 0 ? c.null_count : 0;
  if (level < c.depth) {                 // list node
    p->buffers = {validity, c.offsets[level]};
    a->n_buffers = 2;
    a->n_children = 1;
    p->children.resize(1); p->children[0] = new ArrowArray;
    a->children = p->children.data();
    int64_t child_len = level + 1 < c.n_levels ? c.n_offsets[level + 1] - 1 : c.n_values;
    build_array(a->children[0], c, level + 1, child_len, nullptr);
  } else if (c.elem_type == TFR_T_NULL) {
    a->n_buffers = 0; a->null_count = length;
  } else if (varlen) {
    p->buffers = {validity, c.offsets[c.n_levels - 1], c.values};
    a->n_buffers = 3;
  } else {
    p->buffers = {validity, c.values};
    a->n_buffers = 2;
  }
  a->buffers = p->buffers.data();
}

extern "C" int32_t tfr_batch_export_arrow_host(tfr_batch* b, int32_t column, void* arrow_array, void* arrow_schema) {
  if (!b || !arrow_array || !arrow_schema || column < 0 || column >= (int32_t)b->cols.size()) return fail(TFR_E_INVALID_ARG, "bad argument");
  int32_t rc = tfr_batch_to_host(b, nullptr, (int32_t)b->cols.size());
  if (rc) return rc;
  const tfr_column& c = b->host_cols[column];
  const tfr_schema& S = b->dec->schema;
  std::string nm((const char*)&S.names[S.fields[column].name_off], S.fields[column].name_len);
  build_schema((ArrowSchema*)arrow_schema, nm.c_str(), c.elem_type, c.depth);
  build_array((ArrowArray*)arrow_array, c, 0, c.n_rows, b);
  return TFR_OK;
}
extern "C" int32_t tfr_batch_export_arrow_device(tfr_batch* b, int32_t column, void* arrow_device_array, void* arrow_schema) {
  if (!b || !arrow_device_array || !arrow_schema || column < 0 || column >= (int32_t)b->cols.size()) return fail(TFR_E_INVALID_ARG, "bad argument");
  int32_t rc = tfr_batch_wait(b);
  if (rc) return rc;
  const tfr_column& c = b->cols[column];
  const tfr_schema& S = b->dec->schema;
  std::string nm((const char*)&S.names[S.fields[column].name_off], S.fields[column].name_len);
  build_schema((ArrowSchema*)arrow_schema, nm.c_str(), c.elem_type, c.depth);
  auto* da = (ArrowDeviceArray*)arrow_device_array;
  memset(da, 0, sizeof *da);
  build_array(&da->array, c, 0, c.n_rows, b);
  da->device_id = b->dec->device; da->device_type = 2 /*ARROW_DEVICE_CUDA*/; da->sync_event = nullptr;   // batch already waited
  return TFR_OK;
}

#include "api_encode.inc"
