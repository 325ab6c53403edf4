// infer.cuh -- K6: schema inference.
//
// Replaces TensorFlowInferSchema.apply (M/TensorFlowInferSchema.scala:35-58): rdd.aggregate(empty)(
// inferExampleRowType / inferSequenceExampleRowType, mergeFieldTypes).  Per record every feature of the parsed
// map gets a lattice code (inferField :132-145 + parse*List :147-188: empty list -> null(0), one element ->
// Long/Float/String (1..3), more -> array of it (4..6); FeatureLists: max over the steps, then wrapped into
// ArrayType(ArrayType(T)) (7..9) :98-118); codes of one name merge with findTightestCommonType = max with
// null as identity (:213-228).  Here: warp per record, one map entry per lane (the full-semantics parser of
// decode.cuh), duplicate keys inside a record resolved last-wins before anything is merged, then one
// atomicMax per (name, record) into a device hash table keyed by a 64-bit hash of the name.
#pragma once
#include "common.cuh"
#include "decode.cuh"

struct InferSlot {
  unsigned long long hash;   // 0 = empty
  uint32_t name_off;         // offset of the key bytes inside the batch that first inserted the name
  uint32_t name_len;
  uint32_t code;             // max lattice code seen (0..9)
  uint32_t flags;            // bit0: ArrayType(ArrayType(null)) seen (a FeatureList whose steps are all empty)
};
#define INFER_TABLE_SLOTS 65536u
#define INFER_MAX_ENT 1024     // entries of one map buffered per record for last-wins de-duplication; more raise INF_OVF_ENTRIES
enum { INF_OVF_TABLE = 1u, INF_OVF_ENTRIES = 2u, INF_OVF_KEY = 4u };   // limits hit: reported as an explicit error, never as a silently wrong schema

struct InferArgs {
  const uint8_t* data;
  const uint32_t* rec_off;
  uint32_t n;
  uint32_t verify;
  uint32_t record_type;
  const CrcTables* tabs;
  InferSlot* table;
  uint32_t* first_err;       // [0] min failing record index, [1] INF_OVF_* flags
  uint32_t* status;          // [n]
};

__device__ __forceinline__ unsigned long long hash64(const uint8_t* p, uint32_t n) {
  unsigned long long h = 1469598103934665603ull;
  for (uint32_t i = 0; i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
  return h ? h : 1ull;
}
__device__ __forceinline__ int feat_code(const FeatAcc& a) {       // inferField + parse*List
  if (a.kind == K_NONE) return -1;                                  // RuntimeException("unsupported type ...")
  if (a.n == 0) return 0;
  int base = a.kind == K_INT64 ? 1 : a.kind == K_FLOAT ? 2 : 3;
  return a.n > 1 ? base + 3 : base;
}
__device__ __forceinline__ void infer_merge(InferSlot* table, unsigned long long h, uint32_t name_off, uint32_t name_len, int code, uint32_t* ovf) {
  uint32_t slot = (uint32_t)(h ^ (h >> 32)) & (INFER_TABLE_SLOTS - 1);
  for (uint32_t probe = 0; probe < INFER_TABLE_SLOTS; ++probe) {
    unsigned long long cur = atomicCAS(&table[slot].hash, 0ull, h);
    if (cur == 0ull) { table[slot].name_off = name_off; table[slot].name_len = name_len; cur = h; }
    if (cur == h) {
      if (code == 10) atomicOr(&table[slot].flags, 1u);
      else atomicMax(&table[slot].code, (uint32_t)code);
      return;
    }
    slot = (slot + 1) & (INFER_TABLE_SLOTS - 1);
  }
  atomicOr(ovf, INF_OVF_TABLE);            // more distinct names than slots
}

// one map (Features or FeatureLists) of one record: entries -> (hash, code, key) in shared memory, last wins, merge
__device__ __forceinline__ bool infer_map(const InferArgs& A, Cur body, bool is_flist, unsigned long long* eh, uint32_t* ecode, uint32_t* ekey,
                                          uint32_t& nent, int& err) {
  const uint32_t lane = threadIdx.x & 31;
  // (1) uniform hop collecting entry ranges; 32 at a time parsed one per lane
  uint32_t pend = 0, my_len = 0;
  const uint8_t* my_ptr = nullptr;
  auto flush = [&]() -> bool {
    int code = 0; unsigned long long h = 0; uint32_t koff = 0, klen = 0; bool ok = true;
    if (lane < pend) {
      // entry: last key wins, values merge (same walk as decode.cuh parse_entry, without a schema)
      const uint8_t* key = nullptr;
      Cur c{my_ptr, my_ptr + my_len};
      FeatAcc acc; acc_reset(acc, K_NONE);
      int fl_code = -2;          // -2: no step seen yet
      for (;;) {
        uint32_t tag;
        if (!rd_tag(c, tag)) { ok = false; break; }
        if (tag == 0) break;
        if (tag == 0x0A) {
          uint32_t l; if (!rd_len(c, l) || !utf8_valid(c.p, l)) { ok = false; break; }
          key = c.p; klen = l; c.p += l;
        } else if (tag == 0x12) {
          uint32_t l; if (!rd_len(c, l)) { ok = false; break; }
          Cur v{c.p, c.p + l}; c.p += l;
          if (!is_flist) { if (!feature_scan(v, acc, false, A.data)) { ok = false; break; } }
          else {
            for (;;) {
              uint32_t t2;
              if (!rd_tag(v, t2)) { ok = false; break; }
              if (t2 == 0) break;
              if (t2 != 0x0A) { if (!skip_field(v, t2)) { ok = false; break; } continue; }
              uint32_t sl; if (!rd_len(v, sl)) { ok = false; break; }
              FeatAcc st; acc_reset(st, K_NONE);
              if (!feature_scan(Cur{v.p, v.p + sl}, st, false, A.data)) { ok = false; break; }
              v.p += sl;
              int sc = feat_code(st);
              if (sc < 0) { err = TFR_E_KIND_MISMATCH; }
              if (fl_code == -2) fl_code = sc; else if (sc > fl_code) fl_code = sc;      // reduceLeft(findTightestCommonType)
            }
            if (!ok) break;
          }
        } else if (!skip_field(c, tag)) { ok = false; break; }
      }
      if (ok) {
        h = hash64(key, klen);
        koff = key ? (uint32_t)(key - A.data) : 0;
        if (!is_flist) { code = feat_code(acc); if (code < 0) err = TFR_E_KIND_MISMATCH; }
        else if (fl_code == -2) { err = TFR_E_EMPTY_SCALAR; code = 0; }                 // empty.reduceLeft
        else if (fl_code == 0) code = 10;                                               // ArrayType(ArrayType(null))
        else code = 7 + (fl_code - 1) % 3;                                              // T or [T] -> [[T]]
      }
    }
    if (__any_sync(FULLMASK, !ok)) return false;
    if (lane < pend && klen >= (1u << 24)) atomicOr(&A.first_err[1], INF_OVF_KEY);
    if (lane < pend && nent + lane < INFER_MAX_ENT) {
      eh[nent + lane] = h; ecode[nent + lane] = (uint32_t)(code < 0 ? 0 : code) | (klen << 8); ekey[nent + lane] = koff;
    } else if (lane < pend) {
      // beyond the de-duplication window: last-wins cannot be decided any more
      atomicOr(&A.first_err[1], INF_OVF_ENTRIES);
    }
    nent = min(nent + pend, (uint32_t)INFER_MAX_ENT);
    pend = 0;
    __syncwarp();
    return true;
  };
  for (;;) {
    uint32_t tag;
    if (!rd_tag(body, tag)) return false;
    if (tag == 0) break;
    if (tag != 0x0A) { if (!skip_field(body, tag)) return false; continue; }
    uint32_t l;
    if (!rd_len(body, l)) return false;
    if (lane == pend) { my_ptr = body.p; my_len = l; }
    body.p += l;
    if (++pend == 32 && !flush()) return false;
  }
  if (pend && !flush()) return false;
  return true;
}

__global__ void __launch_bounds__(128) infer_kernel(InferArgs A) {
  extern __shared__ uint32_t smem[];
  uint32_t* stab = smem;
  crc_stage_tables(stab, A.tabs);
  const uint32_t warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wbase = reinterpret_cast<uint8_t*>(smem + CRC_SMEM_WORDS) + (size_t)wid * INFER_MAX_ENT * 16;
  unsigned long long* eh = reinterpret_cast<unsigned long long*>(wbase);
  uint32_t* ecode = reinterpret_cast<uint32_t*>(wbase + INFER_MAX_ENT * 8);
  uint32_t* ekey = reinterpret_cast<uint32_t*>(wbase + INFER_MAX_ENT * 12);
  __syncthreads();
  for (uint32_t row = blockIdx.x * warps + wid; row < A.n; row += gridDim.x * warps) {
    const uint32_t off = A.rec_off[row];
    const uint32_t len = A.rec_off[row + 1] - off - 16;
    const uint8_t* payload = A.data + off + 12;
    int err = 0;
    if (A.verify && crc_mask(crc_warp(stab, payload, len)) != load_u32_unaligned(payload + len)) err = TFR_E_CRC_DATA;
    bool ok = true;
    for (int pass = 0; pass < 2 && ok && !err; ++pass) {       // pass 0: features/context (field 1), pass 1: feature_lists (field 2)
      if (pass == 1 && A.record_type != TFR_RT_SEQUENCE_EXAMPLE) break;
      uint32_t nent = 0;
      Cur top{payload, payload + len};
      for (;;) {
        uint32_t tag;
        if (!rd_tag(top, tag)) { ok = false; break; }
        if (tag == 0) break;
        if (tag == 0x0A || (tag == 0x12 && A.record_type == TFR_RT_SEQUENCE_EXAMPLE)) {
          uint32_t l;
          if (!rd_len(top, l)) { ok = false; break; }
          if ((tag == 0x0A) == (pass == 0) && !infer_map(A, Cur{top.p, top.p + l}, pass == 1, eh, ecode, ekey, nent, err)) { ok = false; break; }
          top.p += l;
        } else if (!skip_field(top, tag)) { ok = false; break; }
      }
      if (!ok) break;
      { uint32_t e = __reduce_max_sync(FULLMASK, err ? (uint32_t)(-err) : 0u); err = e ? -(int)e : 0; }
      if (err) break;
      __syncwarp();
      // Map.put semantics: an entry counts only if no later entry of this map has the same key
      for (uint32_t i = lane; i < nent; i += 32) {
        bool last = true;
        for (uint32_t j = i + 1; j < nent; ++j) if (eh[j] == eh[i]) { last = false; break; }
        if (last) infer_merge(A.table, eh[i], ekey[i], ecode[i] >> 8, (int)(ecode[i] & 0xff), &A.first_err[1]);
      }
      __syncwarp();
    }
    uint32_t st = 0;
    if (err) st = make_status(err, -1);
    else if (!ok) st = make_status(TFR_E_MALFORMED_PROTO, -1);
    if (lane == 0) { A.status[row] = st; if (st) atomicMin(&A.first_err[0], row); }
  }
}

// compact the table: names copied out of the batch, one thread per slot
__global__ void infer_gather_kernel(const InferSlot* __restrict__ table, const uint8_t* __restrict__ data, uint32_t* __restrict__ counters /*[0] entries, [1] bytes*/,
                                    InferSlot* __restrict__ out_entries, uint8_t* __restrict__ out_names, uint32_t names_cap) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= INFER_TABLE_SLOTS || table[s].hash == 0ull) return;
  InferSlot e = table[s];
  uint32_t idx = atomicAdd(&counters[0], 1u);
  uint32_t off = atomicAdd(&counters[1], e.name_len);
  if (off + e.name_len <= names_cap) for (uint32_t i = 0; i < e.name_len; ++i) out_names[off + i] = data[e.name_off + i];
  e.name_off = off;
  out_entries[idx] = e;
}
