// tile.cuh -- fast path of the decode: shared-memory record tiles, one record per thread.
//
// Same reference functions as decode.cuh (TFRecordReader payload CRC, Example.parseFrom,
// deserializeExample: M/TFRecordFileReader.scala:49-81, M/TFRecordDeserializer.scala:21-35,68-124), for
// the shape of data the reference writer itself produces ("canonical"):
//   payload  = 0A len Features                       (exactly one field)
//   Features = { 0A elen  0A klen key  12 vlen Feature }*
//   Feature  = kindtag llen List ;  Int64List/FloatList = [0A plen packed] ; BytesList = { 0A blen bytes }*
// Everything else that protobuf allows (unknown fields, merges, unpacked encodings, duplicate keys,
// overlong tags ...), every semantic error and every CRC mismatch is DETECTED here and makes the kernel
// raise a flag; the host then re-runs the batch through the general kernels of decode.cuh, which
// implement the full semantics.  So the fast path never changes a result, it only skips work.
//
// Mapping: one CTA = one tile = 32 consecutive records (rows 32t .. 32t+31, contiguous in memory),
// TILE_PARSE_WARPS parse warps + 1 CRC warp, all working on the same staged bytes.
//   1. one thread arms an mbarrier and issues cp.async.bulk (TMA bulk copy, SASS UBLKCP) of the tile's
//      byte range into shared memory; meanwhile all threads stage the slicing-by-8 CRC tables and the
//      schema (field table + names) into shared memory.
//   2. role split over the same staged bytes: the last warp computes the masked CRC-32C of record `lane`
//      (serial slicing-by-8, 8 bytes per step); parse warp w owns the map entries with index = w mod W of
//      record `lane`: it fully parses those and only hops over the others (`0A elen` -> p += elen).  An
//      owned entry is first matched against a per-field TEMPLATE of its constant bytes
//      (0A ? 0A klen key 12 ? kind ?: a few masked word compares); only when that fails is it parsed
//      byte by byte (hash lookup of the key, full checks).  Shared-memory capacity limits how many
//      records an SM can stage, so more dependent-chains per staged record = more warps to hide latency.
//   3. lane = row, rows are 32-aligned: every column store of the warp covers 32 consecutive rows
//      (coalesced by construction, no transpose) and validity bitmaps are one __ballot_sync per field.
//   4. variable-width columns either write element counts + source offsets (scan + decode_pass2_kernel
//      finish them), or -- once the decoder has learned that every such column has a uniform shape
//      (FloatList[8], 16-byte BytesList ...) -- write the values at row * L in the same pass and only
//      verify the shape ("uniform-shape speculation": input read once, output written once).
#pragma once
#include "common.cuh"
#include "decode.cuh"

#define TILE_ROWS 32
#ifndef TILE_PARSE_WARPS
#define TILE_PARSE_WARPS 8     // power of two
#endif
#define TILE_THREADS ((TILE_PARSE_WARPS + 1) * 32)
#define TILE_TPL_WORDS 5          // template covers up to 20 bytes: key names up to 12 bytes

// constant bytes of a canonical map entry of one schema field, for masked word compares
struct FieldTemplate {
  uint32_t words[TILE_TPL_WORDS];
  uint32_t mask[TILE_TPL_WORDS];
  uint32_t n_words;             // 0: no template (long name): generic parse
  uint32_t kind;                // K_*
};

struct TileArgs {
  const uint8_t* data;
  uint32_t nbytes;
  const uint32_t* rec_off;      // [n+1]
  uint32_t n;                   // rows in the batch (stride of the scratch arrays)
  uint32_t tile_cap;            // bytes of shared memory reserved for the record bytes of one tile
  uint32_t verify;
  uint32_t names_bytes;
  DevSchema sch;
  const FieldTemplate* templates;   // [n_fields]
  const CrcTables* tabs;
  uint8_t* bitmaps;             // [nf][nb_stride] Arrow validity bitmaps, written directly
  uint32_t nb_stride;
  unsigned long long* null_counts;   // [nf]
  void* const* fix_values;      // [n_fix]
  uint32_t* cnt;                // [n_cnt][n]   (count mode)
  uint32_t* src;                // [n_var][n]
  uint8_t* cflag;               // [n_var][n]
  const int32_t* uniform_len;   // [n_var] >= 0: speculated per-row count (elements, or bytes for scalar string/binary); -1: count mode
  void* const* var_values;      // [n_var] leaf buffers (uniform mode)
  uint32_t* flags;              // [0] bit0: fall back to the general path, bit1: a uniform-shape speculation failed
};

enum { TF_FALLBACK = 1u, TF_SHAPE = 2u };

// ---- mbarrier + bulk async copy (PTX; sm_90+) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!ok);
}

// ---- shared-memory byte helpers (offsets into the tile, not pointers: 32-bit address arithmetic) --
// The loads are explicit ld.shared on a 32-bit shared-window address held in a register: through a generic pointer the
// compiler re-derives the window base (S2R SR_CgaCtaId + LEA) at most access sites, ~5% of the kernel's instructions.
// The tile is read-only after the mbarrier wait; `s` is produced by a volatile asm placed after that wait so that no
// load can be scheduled above it.
struct Tile {
  const uint8_t* b;   // tile base in shared memory (16-byte aligned), generic pointer for the rare helper calls
  uint32_t s;         // the same address in the shared window
  __device__ __forceinline__ uint32_t u8(uint32_t o) const {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(s + o));
    return v;
  }
  __device__ __forceinline__ int32_t i8(uint32_t o) const {
    int32_t v;
    asm("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(s + o));
    return v;
  }
  __device__ __forceinline__ uint32_t w32(uint32_t o) const {   // aligned word
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(s + o));
    return v;
  }
};
__device__ __forceinline__ uint32_t t_u32(const Tile& t, uint32_t o) {   // 4 bytes at any alignment
  const uint32_t a = o & ~3u;
  uint32_t sh = (o & 3u) * 8;
  uint32_t lo = t.w32(a);
  return sh ? __funnelshift_r(lo, t.w32(a + 4), sh) : lo;
}
// length varint: 1..5 bytes (minimal or not); false if it runs past `end` or is longer (-> general path)
__device__ __forceinline__ bool t_len(const Tile& t, uint32_t& p, uint32_t end, uint32_t& v) {
  if (p >= end) return false;
  uint32_t b = t.u8(p++);
  if (b < 0x80) { v = b; return true; }
  uint32_t r = b & 0x7f;
#pragma unroll 1
  for (int sh = 7; sh < 35; sh += 7) {
    if (p >= end) return false;
    b = t.u8(p++);
    r |= (b & 0x7f) << sh;
    if (b < 0x80) { v = r; return (int32_t)r >= 0; }
  }
  return false;
}

// ---- per-thread CRC-32C over shared memory: slicing-by-8 with instruction-level parallelism ----
// One 8-byte step of the register: s8 = 8 tables of 256 words.
__device__ __forceinline__ uint32_t crc_fold8(const uint32_t* s8, uint32_t c, uint32_t lo, uint32_t hi) {
  uint32_t a = lo ^ c;
  return s8[7 * 256 + (a & 0xff)] ^ s8[6 * 256 + ((a >> 8) & 0xff)] ^ s8[5 * 256 + ((a >> 16) & 0xff)] ^ s8[4 * 256 + (a >> 24)] ^
         s8[3 * 256 + (hi & 0xff)] ^ s8[2 * 256 + ((hi >> 8) & 0xff)] ^ s8[1 * 256 + ((hi >> 16) & 0xff)] ^ s8[(hi >> 24)];
}
// serial: state `c` over n bytes at tile offset o (any alignment)
__device__ __forceinline__ uint32_t crc_serial_s8(const uint32_t* s8, const uint8_t* base, uint32_t o, uint32_t n, uint32_t c) {
  while (n && (o & 3)) { c = (c >> 8) ^ s8[(c ^ base[o++]) & 0xff]; --n; }
  const uint32_t* w = reinterpret_cast<const uint32_t*>(base + o);
  uint32_t nw = n >> 3;
#pragma unroll 2
  for (uint32_t i = 0; i < nw; ++i) c = crc_fold8(s8, c, w[2 * i], w[2 * i + 1]);
  o += nw * 8; n &= 7;
  while (n--) c = (c >> 8) ^ s8[(c ^ base[o++]) & 0xff];
  return c;
}
// A serial CRC is one long dependent chain (one table-lookup round trip per 8 bytes).  The payload is cut
// into 512-byte segments; three segment chains run interleaved in the same thread (independent
// registers -> the LDS latencies overlap), each segment state is shifted over the segments that follow
// it with ONE GF(2) multiply by the table constant x^(8*512*m) (xp512, in shared memory) and the
// < 512-byte tail is folded serially from the combined state.  CRC(A||B) = CRC_B(0) ^ shift_|B|(CRC_A).
#define CRC_SEG 512u
__device__ __forceinline__ uint32_t crc_segmented(const uint32_t* s8, const uint32_t* xp, const uint8_t* base, uint32_t o, uint32_t n) {
  const uint32_t nF = n / CRC_SEG;
  if (nF == 0 || nF > 127) return ~crc_serial_s8(s8, base, o, n, 0xFFFFFFFFu);
  const uint32_t sh = (o & 3u) * 8;
  const uint32_t* W = reinterpret_cast<const uint32_t*>(base + (o & ~3u));     // aligned words; chain k starts at W + k*128
  uint32_t acc = 0, j = 0;
  for (; j + 3 <= nF; j += 3) {
    const uint32_t* w0 = W + j * (CRC_SEG / 4);
    const uint32_t* w1 = w0 + CRC_SEG / 4;
    const uint32_t* w2 = w1 + CRC_SEG / 4;
    uint32_t c0 = j == 0 ? 0xFFFFFFFFu : 0u, c1 = 0, c2 = 0;
    uint32_t k0 = w0[0], k1 = w1[0], k2 = w2[0];                               // carry words for the funnel shifts
#pragma unroll 4
    for (uint32_t i = 0; i < CRC_SEG / 8; ++i) {
      uint32_t a0 = w0[2 * i + 1], b0 = w0[2 * i + 2], a1 = w1[2 * i + 1], b1 = w1[2 * i + 2], a2 = w2[2 * i + 1], b2 = w2[2 * i + 2];
      c0 = crc_fold8(s8, c0, __funnelshift_r(k0, a0, sh), __funnelshift_r(a0, b0, sh)); k0 = b0;
      c1 = crc_fold8(s8, c1, __funnelshift_r(k1, a1, sh), __funnelshift_r(a1, b1, sh)); k1 = b1;
      c2 = crc_fold8(s8, c2, __funnelshift_r(k2, a2, sh), __funnelshift_r(a2, b2, sh)); k2 = b2;
    }
    acc ^= gf2_mulmod(xp[nF - 1 - j], c0) ^ gf2_mulmod(xp[nF - 2 - j], c1) ^ gf2_mulmod(xp[nF - 3 - j], c2);
  }
  for (; j < nF; ++j) {
    uint32_t c = crc_serial_s8(s8, base, o + j * CRC_SEG, CRC_SEG, j == 0 ? 0xFFFFFFFFu : 0u);
    acc ^= gf2_mulmod(xp[nF - 1 - j], c);
  }
  return ~crc_serial_s8(s8, base, o + nF * CRC_SEG, n - nF * CRC_SEG, acc);
}

// shared memory layout (dynamic), all sections 16-byte aligned:
//   [0,16) mbarrier | CRC tables 8 KiB + xp512 512 B | seen masks [W][32][2] u64 | DevField[nf] | FieldTemplate[nf] | names | tile bytes
#define TILE_SEEN_BYTES (TILE_PARSE_WARPS * 32 * 16)
__host__ __device__ inline uint32_t tile_schema_smem(uint32_t nf, uint32_t names_bytes) {
  return ((nf * (uint32_t)sizeof(DevField) + 15u) & ~15u) + ((nf * (uint32_t)sizeof(FieldTemplate) + 15u) & ~15u) + ((names_bytes + 15u) & ~15u);
}
__host__ __device__ inline uint32_t tile_smem_bytes(uint32_t nf, uint32_t names_bytes, uint32_t tile_cap) {
  return 16 + 8192 + 512 + TILE_SEEN_BYTES + tile_schema_smem(nf, names_bytes) + tile_cap + 64;   // +64: template compares may look a few bytes past the tile
}

template <bool SEQ>
__global__ void __launch_bounds__(TILE_THREADS) decode_tile_kernel(TileArgs A) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* s8 = reinterpret_cast<uint32_t*>(smem_raw + 16);                               // 8 KiB
  unsigned long long* sseen = reinterpret_cast<unsigned long long*>(smem_raw + 16 + 8192 + 512);
  const uint32_t nf = (uint32_t)A.sch.n_fields;
  uint8_t* sbase = smem_raw + 16 + 8192 + 512 + TILE_SEEN_BYTES;
  DevField* sfields = reinterpret_cast<DevField*>(sbase);
  FieldTemplate* stpl = reinterpret_cast<FieldTemplate*>(sbase + ((nf * (uint32_t)sizeof(DevField) + 15u) & ~15u));
  uint8_t* snames = sbase + ((nf * (uint32_t)sizeof(DevField) + 15u) & ~15u) + ((nf * (uint32_t)sizeof(FieldTemplate) + 15u) & ~15u);
  uint8_t* tile_b = sbase + tile_schema_smem(nf, A.names_bytes);
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;                          // warps 0..W-1 parse, warp W = CRC

  const uint32_t row0 = blockIdx.x * TILE_ROWS;
  const uint32_t rows = min((uint32_t)TILE_ROWS, A.n - row0);
  const uint32_t first = A.rec_off[row0], last = A.rec_off[row0 + rows];
  const uint32_t g0 = first & ~15u;
  const uint32_t span = last - g0;
  if (span > A.tile_cap) {                                        // records too large for the tile: general path
    if (threadIdx.x == 0) atomicOr(A.flags, TF_FALLBACK);
    return;
  }
  const uint32_t copy_bytes = (span + 15u) & ~15u;                // may read < 16 bytes past the end (input buffers are padded)
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, copy_bytes);
    uint32_t done = 0;
    while (done < copy_bytes) {
      uint32_t part = min(copy_bytes - done, 32768u);
      bulk_g2s(tile_b + done, A.data + g0 + done, part, bar);
      done += part;
    }
  }
  {   // meanwhile: CRC tables + schema into shared memory
    const uint32_t* g = reinterpret_cast<const uint32_t*>(A.tabs->s8);
    for (uint32_t i = threadIdx.x; i < 2048 + 128; i += TILE_THREADS) s8[i] = g[i];      // s8 tables, then xp512 (contiguous in CrcTables)
    const uint32_t* gf = reinterpret_cast<const uint32_t*>(A.sch.fields);
    uint32_t* sf = reinterpret_cast<uint32_t*>(sfields);
    for (uint32_t i = threadIdx.x; i < nf * (uint32_t)(sizeof(DevField) / 4); i += TILE_THREADS) sf[i] = gf[i];
    const uint32_t* gt = reinterpret_cast<const uint32_t*>(A.templates);
    uint32_t* stw = reinterpret_cast<uint32_t*>(stpl);
    for (uint32_t i = threadIdx.x; i < nf * (uint32_t)(sizeof(FieldTemplate) / 4); i += TILE_THREADS) stw[i] = gt[i];
    for (uint32_t i = threadIdx.x; i < A.names_bytes; i += TILE_THREADS) snames[i] = A.sch.names[i];
  }
  __syncthreads();
  mbar_wait(bar, 0);

  const bool active = lane < rows;
  const uint32_t row = row0 + lane;
  uint32_t off = 0, len = 0;
  if (active) { off = A.rec_off[row]; len = A.rec_off[row + 1] - off - 16; }
  const uint32_t pay = off - g0 + 12;            // payload offset inside the tile
  const uint32_t end = pay + len;
  Tile T;
  T.b = tile_b;
  asm volatile("mov.u32 %0, %1;" : "=r"(T.s) : "r"(smem_u32(tile_b)) : "memory");   // ordered after mbar_wait

  // =============================== last warp: CRC ===============================
  if (wid == TILE_PARSE_WARPS) {
    if (active && A.verify) {
      // the frame index chained the headers without checking them on the fast path: check the length CRC here
      if (crc_mask(crc_u64(s8, t_u32(T, pay - 12), t_u32(T, pay - 8))) != t_u32(T, pay - 4)) atomicOr(A.flags, TF_FALLBACK);
      uint32_t crc = crc_segmented(s8, s8 + 2048, tile_b, pay, len);
      if (crc_mask(crc) != t_u32(T, end)) atomicOr(A.flags, TF_FALLBACK);      // the general path reports the error at the right record
    }
    return;
  }

  // =============================== warps 0..W-1: parse ===============================
  bool bad = false;
  uint32_t entry_idx = 0;
  uint32_t shape_bad = 0;
  unsigned long long seen_lo = 0, seen_hi = 0;
  if (active) {
    uint32_t p = pay, L = 0;
    uint32_t cend = end, fl_start = end, fl_end = end;      // context/features region = [p, cend), feature_lists = [fl_start, fl_end)
    // Example { features = 1 } / SequenceExample { context = 1, feature_lists = 2 }: exactly these fields, in this order
    if (len < 2 || T.u8(p) != 0x0A) bad = true;
    else {
      ++p;
      if (!t_len(T, p, end, L) || end - p < L) bad = true;
      else if (!SEQ) { if (p + L != end) bad = true; }
      else {
        cend = p + L;
        uint32_t q = cend, L2 = 0;
        if (q >= end || T.u8(q) != 0x12) bad = true;
        else { ++q; if (!t_len(T, q, end, L2) || q + L2 != end) bad = true; else { fl_start = q; fl_end = end; } }
      }
    }
    uint32_t next_f = wid;                 // in-order data: this warp's k-th owned entry is field wid + k*W
    while (!bad && p < cend) {
      // hop over the entries the other parse warps own (`0A elen ...`): a tight loop, one byte load + add per entry.
      // The owner validates those entries; p + 1 <= cend is always inside the tile and an overshoot is caught by the
      // p == cend check after the loop.
      uint32_t skip = (wid - entry_idx) & (TILE_PARSE_WARPS - 1);
      entry_idx += skip;
      for (bool wide = true; wide;) {
        wide = false;
        while (skip && p < cend) {                           // the tight part: single-byte entry lengths only
          const int32_t b1 = T.i8(p + 1);
          if (b1 < 0) { wide = true; break; }
          p += 2u + (uint32_t)b1;
          --skip;
        }
        if (wide) {                                          // an entry of 128+ bytes: full length varint, then back to the loop
          uint32_t q = p + 1, el;
          if (!t_len(T, q, cend, el)) { bad = true; break; }
          p = q + el;
          --skip;
        }
      }
      entry_idx -= skip;                                     // hops not performed (ran into cend)
      if (bad || p >= cend) break;
      ++entry_idx;                                           // the entry this warp owns
      // ---- owned entry: try the expected field's template first ----
      uint32_t eend, vlen;
      int f = -1;
      bool located = false;
      if (next_f < nf && stpl[next_f].n_words) {
        const FieldTemplate& tp = stpl[next_f];
        const uint32_t klen = sfields[next_f].name_len;
        uint32_t diff = 0;
#pragma unroll
        for (int w = 0; w < TILE_TPL_WORDS; ++w)
          if ((uint32_t)w < tp.n_words) diff |= (t_u32(T, p + 4 * w) ^ tp.words[w]) & tp.mask[w];
        const uint32_t elen = T.u8(p + 1), vl = T.u8(p + 5 + klen), ll = T.u8(p + 7 + klen);
        // single-byte lengths that nest exactly: entry = key part (klen+2) + 2 + value; value = 2 + list
        if (diff == 0 && elen < 0x80 && elen == klen + 4 + vl && vl == ll + 2 && p + 2 + elen <= cend) {
          f = (int)next_f; eend = p + 2 + elen; vlen = vl; located = true;
          p += klen + 6;                                      // at the kind tag
        }
      }
      if (!located) {
        // ---- generic: 0A elen 0A klen key 12 vlen, key looked up by hash ----
        uint32_t elen, klen;
        if (T.u8(p) != 0x0A) { bad = true; break; }
        ++p;
        if (!t_len(T, p, cend, elen) || cend - p < elen) { bad = true; break; }
        eend = p + elen;
        if (p >= eend || T.u8(p) != 0x0A) { bad = true; break; }
        ++p;
        if (!t_len(T, p, eend, klen) || eend - p < klen) { bad = true; break; }
        const uint32_t key = p;
        p += klen;
        if (p >= eend || T.u8(p) != 0x12) { bad = true; break; }
        ++p;
        if (!t_len(T, p, eend, vlen) || p + vlen != eend) { bad = true; break; }
        uint32_t hi = 0;
        for (uint32_t i = 0; i < klen; ++i) hi |= T.u8(key + i);
        if (hi >= 0x80 && !utf8_valid(T.b + key, klen)) { bad = true; break; }   // malformed key: the general path reports it
        uint32_t h = name_hash(T.b + key, klen);
        uint32_t slot = h & (uint32_t)A.sch.ht_mask;
        for (;;) {
          int cand = A.sch.ht[slot];
          if (cand < 0) break;
          if (sfields[cand].hash == h && sfields[cand].name_len == klen) {
            const uint8_t* nm = snames + sfields[cand].name_off;
            uint32_t diff = 0;
            for (uint32_t i = 0; i < klen; ++i) diff |= T.u8(key + i) ^ nm[i];
            if (diff == 0) { f = cand; break; }
          }
          slot = (slot + 1) & (uint32_t)A.sch.ht_mask;
        }
      }
      if (f >= 0 && sfields[f].elem_type == TFR_T_NULL) f = -1;                 // NullType: always null, value only validated
      if (f >= 0) {
        unsigned long long bit = 1ull << (f & 63);
        if (f < 64) { if (seen_lo & bit) { bad = true; break; } seen_lo |= bit; }    // duplicate key (inside this warp's entries)
        else { if (seen_hi & bit) { bad = true; break; } seen_hi |= bit; }
        next_f = (uint32_t)f + TILE_PARSE_WARPS;
      }
      // ---- Feature: exactly one oneof member spanning the value ----
      if (vlen == 0) { if (f >= 0) bad = true; p = eend; continue; }              // kind not set: an error if the schema wants it
      uint32_t kt = T.u8(p++), llen;
      uint32_t kind = kt == 0x0A ? K_BYTES : kt == 0x12 ? K_FLOAT : kt == 0x1A ? K_INT64 : K_NONE;
      if (kind == K_NONE || !t_len(T, p, eend, llen) || p + llen != eend) { bad = true; break; }
      const DevField* fd = f >= 0 ? &sfields[f] : nullptr;
      if (fd && ((uint32_t)fd->kind != kind || fd->depth > 1)) { bad = true; break; }   // kind mismatch / nesting: error path
      if (kind == K_BYTES) {
        // BytesList: { 0A blen bytes }*
        uint32_t n = 0, first_off = 0, first_len = 0, first_data = 0, total = 0;
        const uint32_t body = p;
        const bool is_str = fd && fd->elem_type == TFR_T_STRING;
        while (p < eend) {
          uint32_t bl;
          if (T.u8(p) != 0x0A) { bad = true; break; }
          ++p;
          const uint32_t lp = p;
          if (!t_len(T, p, eend, bl) || eend - p < bl) { bad = true; break; }
          if (is_str) {
            // StringType = Java UTF-8 decode/re-encode: identity for well-formed input; malformed input needs the
            // U+FFFD transcode, which only the general path implements
            uint32_t acc = 0;
            for (uint32_t i = 0; i < bl; ++i) acc |= T.u8(p + i);
            if (acc >= 0x80 && !utf8_valid(T.b + p, bl)) { bad = true; break; }
          }
          if (n == 0) { first_off = lp + g0; first_len = bl; first_data = p; }
          ++n; total += bl;
          p += bl;
        }
        if (bad) break;
        if (fd) {
          if (fd->depth == 0) {
            if (n == 0) { bad = true; break; }                                  // .head of an empty list: error path
            const int32_t ul = A.uniform_len[fd->var_slot];
            if (ul >= 0) {
              if ((uint32_t)ul != first_len) shape_bad = 1;
              else {
                uint8_t* dst = reinterpret_cast<uint8_t*>(A.var_values[fd->var_slot]) + (size_t)row * (uint32_t)ul;
                if ((ul & 15) == 0) {
                  for (uint32_t i = 0; i < (uint32_t)ul; i += 16) {
                    uint4 v;
                    v.x = t_u32(T, first_data + i); v.y = t_u32(T, first_data + i + 4); v.z = t_u32(T, first_data + i + 8); v.w = t_u32(T, first_data + i + 12);
                    *reinterpret_cast<uint4*>(dst + i) = v;
                  }
                } else for (uint32_t i = 0; i < (uint32_t)ul; ++i) dst[i] = T.u8(first_data + i);
              }
            } else {
              A.cnt[(size_t)fd->cnt_slot * A.n + row] = first_len;
              A.src[(size_t)fd->var_slot * A.n + row] = first_off;
              A.cflag[(size_t)fd->var_slot * A.n + row] = CF_CANON;
            }
          } else {
            // ArrayType(String/Binary): two offset levels -> always count mode (never speculated)
            A.cnt[(size_t)fd->cnt_slot * A.n + row] = n;
            A.cnt[(size_t)(fd->cnt_slot + 1) * A.n + row] = total;
            A.src[(size_t)fd->var_slot * A.n + row] = body + g0;
            A.cflag[(size_t)fd->var_slot * A.n + row] = CF_CANON;
          }
        }
      } else {
        // Int64List / FloatList: empty, or one packed field spanning the list
        uint32_t plen = 0, pk = p;
        if (llen != 0) {
          if (T.u8(p) != 0x0A) { bad = true; break; }
          ++p;
          if (!t_len(T, p, eend, plen) || p + plen != eend) { bad = true; break; }
          pk = p;
        }
        uint32_t n;
        uint64_t v0 = 0;
        if (kind == K_FLOAT) {
          if (plen & 3) { bad = true; break; }
          n = plen >> 2;
        } else {
          // varints: count terminators, every run <= 10 bytes, the last byte terminates; first value decoded on the way
          n = 0;
          uint32_t run = 0;
          if (plen == 1) {                                   // the most common case: one small value
            v0 = T.u8(pk);
            if (v0 & 0x80) { bad = true; break; }
            n = 1;
          } else {
            for (uint32_t i = 0; i < plen; ++i) {
              uint32_t b = T.u8(pk + i);
              if (n == 0) v0 |= (uint64_t)(b & 0x7f) << (7 * run);
              if (b & 0x80) { if (++run >= 10) { bad = true; break; } } else { ++n; run = 0; }
            }
            if (bad || run) { bad = true; break; }
          }
        }
        if (fd) {
          if (fd->depth == 0) {
            if (n == 0) { bad = true; break; }
            void* vp = A.fix_values[fd->fix_slot];
            if (kind == K_FLOAT) {
              uint32_t bits = t_u32(T, pk);
              if (fd->elem_type == TFR_T_FLOAT32) reinterpret_cast<uint32_t*>(vp)[row] = bits;
              else reinterpret_cast<double*>(vp)[row] = (double)__uint_as_float(bits);
            } else {
              if (fd->elem_type == TFR_T_INT64) reinterpret_cast<int64_t*>(vp)[row] = (int64_t)v0;
              else reinterpret_cast<int32_t*>(vp)[row] = (int32_t)(uint32_t)v0;
            }
          } else {
            const int32_t ul = A.uniform_len[fd->var_slot];
            if (ul >= 0) {
              if ((uint32_t)ul != n) shape_bad = 1;
              else if (kind == K_FLOAT) {
                if (fd->elem_type == TFR_T_FLOAT32) {
                  uint32_t* dst = reinterpret_cast<uint32_t*>(A.var_values[fd->var_slot]) + (size_t)row * n;
                  if ((n & 3) == 0) {
                    for (uint32_t i = 0; i < n; i += 4) {
                      uint4 v;
                      v.x = t_u32(T, pk + 4 * i); v.y = t_u32(T, pk + 4 * i + 4); v.z = t_u32(T, pk + 4 * i + 8); v.w = t_u32(T, pk + 4 * i + 12);
                      *reinterpret_cast<uint4*>(dst + i) = v;
                    }
                  } else for (uint32_t i = 0; i < n; ++i) dst[i] = t_u32(T, pk + 4 * i);
                } else {
                  double* dst = reinterpret_cast<double*>(A.var_values[fd->var_slot]) + (size_t)row * n;
                  for (uint32_t i = 0; i < n; ++i) dst[i] = (double)__uint_as_float(t_u32(T, pk + 4 * i));
                }
              } else {
                uint32_t q = pk;
                for (uint32_t e = 0; e < n; ++e) {
                  uint64_t v = 0; uint32_t sh = 0;
                  for (;;) { uint32_t b = T.u8(q++); v |= (uint64_t)(b & 0x7f) << sh; sh += 7; if (b < 0x80) break; }
                  if (fd->elem_type == TFR_T_INT64) reinterpret_cast<int64_t*>(A.var_values[fd->var_slot])[(size_t)row * n + e] = (int64_t)v;
                  else reinterpret_cast<int32_t*>(A.var_values[fd->var_slot])[(size_t)row * n + e] = (int32_t)(uint32_t)v;
                }
              }
            } else {
              A.cnt[(size_t)fd->cnt_slot * A.n + row] = n;
              A.src[(size_t)fd->var_slot * A.n + row] = pk + g0;
              A.cflag[(size_t)fd->var_slot * A.n + row] = CF_CANON;
            }
          }
        }
      }
      p = eend;
    }
    if (p != cend) bad = true;
    // ---- SequenceExample.feature_lists: { 0A elen 0A klen key 12 vlen FeatureList }*, FeatureList = { 0A flen Feature }* ----
    p = fl_start;
    while (SEQ && !bad && p < fl_end) {
      uint32_t skip = (wid - entry_idx) & (TILE_PARSE_WARPS - 1);
      while (skip && p < fl_end) {
        ++entry_idx;
        const uint32_t b1 = T.u8(p + 1);
        if (b1 < 0x80) p += 2 + b1;
        else {
          uint32_t q = p + 1, el;
          if (!t_len(T, q, fl_end, el)) { bad = true; break; }
          p = q + el;
        }
        --skip;
      }
      if (bad || p >= fl_end) break;
      ++entry_idx;                                           // the entry this warp owns
      uint32_t elen, klen, vlen;
      if (T.u8(p) != 0x0A) { bad = true; break; }
      ++p;
      const uint32_t entry_pos = p;                           // the entry's length varint: what pass 2's FeatureList walker starts from
      if (!t_len(T, p, fl_end, elen) || fl_end - p < elen) { bad = true; break; }
      const uint32_t eend = p + elen;
      if (p >= eend || T.u8(p) != 0x0A) { bad = true; break; }
      ++p;
      if (!t_len(T, p, eend, klen) || eend - p < klen) { bad = true; break; }
      const uint32_t key = p;
      p += klen;
      if (p >= eend || T.u8(p) != 0x12) { bad = true; break; }
      ++p;
      if (!t_len(T, p, eend, vlen) || p + vlen != eend) { bad = true; break; }
      uint32_t hi = 0;
      for (uint32_t i = 0; i < klen; ++i) hi |= T.u8(key + i);
      if (hi >= 0x80 && !utf8_valid(T.b + key, klen)) { bad = true; break; }
      int f = -1;
      {
        uint32_t h = name_hash(T.b + key, klen);
        uint32_t slot = h & (uint32_t)A.sch.ht_mask;
        for (;;) {
          int cand = A.sch.ht[slot];
          if (cand < 0) break;
          if (sfields[cand].hash == h && sfields[cand].name_len == klen) {
            const uint8_t* nm = snames + sfields[cand].name_off;
            uint32_t diff = 0;
            for (uint32_t i = 0; i < klen; ++i) diff |= T.u8(key + i) ^ nm[i];
            if (diff == 0) { f = cand; break; }
          }
          slot = (slot + 1) & (uint32_t)A.sch.ht_mask;
        }
      }
      if (f >= 0 && sfields[f].elem_type == TFR_T_NULL) f = -1;
      const DevField* fd = f >= 0 ? &sfields[f] : nullptr;
      if (fd) {
        if (fd->depth != 2) { bad = true; break; }            // array of heads / scalar from a FeatureList: general path
        unsigned long long bit = 1ull << (f & 63);
        // a name present in context AND feature_lists (context wins) or twice here: general path
        if (f < 64) { if (seen_lo & bit) { bad = true; break; } seen_lo |= bit; }
        else { if (seen_hi & bit) { bad = true; break; } seen_hi |= bit; }
      }
      uint32_t steps = 0, tot_n = 0, tot_bytes = 0;
      while (p < eend) {                                      // steps
        uint32_t flen;
        if (T.u8(p) != 0x0A) { bad = true; break; }
        ++p;
        if (!t_len(T, p, eend, flen) || eend - p < flen) { bad = true; break; }
        const uint32_t fend = p + flen;
        ++steps;
        if (flen == 0) { if (fd) bad = true; continue; }     // kind not set: an error if the schema wants the column
        uint32_t kt = T.u8(p++), llen;
        uint32_t kind = kt == 0x0A ? K_BYTES : kt == 0x12 ? K_FLOAT : kt == 0x1A ? K_INT64 : K_NONE;
        if (kind == K_NONE || !t_len(T, p, fend, llen) || p + llen != fend) { bad = true; break; }
        if (fd && (uint32_t)fd->kind != kind) { bad = true; break; }
        if (kind == K_BYTES) {
          const bool is_str = fd && fd->elem_type == TFR_T_STRING;
          while (p < fend) {
            uint32_t bl;
            if (T.u8(p) != 0x0A) { bad = true; break; }
            ++p;
            if (!t_len(T, p, fend, bl) || fend - p < bl) { bad = true; break; }
            if (is_str) {
              uint32_t acc = 0;
              for (uint32_t i = 0; i < bl; ++i) acc |= T.u8(p + i);
              if (acc >= 0x80 && !utf8_valid(T.b + p, bl)) { bad = true; break; }
            }
            ++tot_n; tot_bytes += bl;
            p += bl;
          }
          if (bad) break;
        } else if (llen != 0) {
          uint32_t plen;
          if (T.u8(p) != 0x0A) { bad = true; break; }
          ++p;
          if (!t_len(T, p, fend, plen) || p + plen != fend) { bad = true; break; }
          if (kind == K_FLOAT) {
            if (plen & 3) { bad = true; break; }
            tot_n += plen >> 2;
          } else {
            uint32_t run = 0;
            for (uint32_t i = 0; i < plen; ++i) {
              if (T.u8(p + i) & 0x80) { if (++run >= 10) { bad = true; break; } } else { ++tot_n; run = 0; }
            }
            if (bad || run) { bad = true; break; }
          }
        }
        p = fend;
      }
      if (bad) break;
      if (fd) {
        A.cnt[(size_t)fd->cnt_slot * A.n + row] = steps;
        A.cnt[(size_t)(fd->cnt_slot + 1) * A.n + row] = tot_n;
        if (fd->n_levels == 3) A.cnt[(size_t)(fd->cnt_slot + 2) * A.n + row] = tot_bytes;
        A.src[(size_t)fd->var_slot * A.n + row] = entry_pos + g0;
        A.cflag[(size_t)fd->var_slot * A.n + row] = CF_FLIST;
      }
      p = eend;
    }
    if (!bad && p != fl_end) bad = true;
  }
  // ---- merge the parse warps' seen masks (a key seen by two warps is a duplicate) ----
  sseen[(wid * 32 + lane) * 2] = seen_lo;
  sseen[(wid * 32 + lane) * 2 + 1] = seen_hi;
  asm volatile("bar.sync 1, %0;" ::"r"(TILE_PARSE_WARPS * 32) : "memory");
  if (wid == 0) {
    unsigned long long all_lo = 0, all_hi = 0;
#pragma unroll
    for (int w = 0; w < TILE_PARSE_WARPS; ++w) {
      unsigned long long a = sseen[(w * 32 + lane) * 2], b = sseen[(w * 32 + lane) * 2 + 1];
      if ((all_lo & a) | (all_hi & b)) bad = true;                            // duplicate key across warps: last-wins -> general path
      all_lo |= a; all_hi |= b;
    }
    seen_lo = all_lo; seen_hi = all_hi;
    // ---- validity bitmaps by ballot (rows are 32-aligned), absent fields -> null / error ----
    const uint32_t act_mask = __ballot_sync(FULLMASK, active);
    for (uint32_t f0 = 0; f0 < nf; f0 += 32) {
      uint32_t my_word = 0;
      const uint32_t lim = min(32u, nf - f0);
      for (uint32_t j = 0; j < lim; ++j) {
        const uint32_t f = f0 + j;
        const bool present = f < 64 ? (seen_lo >> f) & 1 : (seen_hi >> (f - 64)) & 1;
        const uint32_t m = __ballot_sync(FULLMASK, present && active);
        if (lane == j) my_word = m;
        if (active && !present && sfields[f].elem_type != TFR_T_NULL) {
          const DevField& fd = sfields[f];
          if (!fd.nullable) bad = true;                                       // NullPointerException: error path
          if (fd.fix_slot >= 0) {
            void* vp = A.fix_values[fd.fix_slot];
            if (fd.width == 8) reinterpret_cast<uint64_t*>(vp)[row] = 0; else reinterpret_cast<uint32_t*>(vp)[row] = 0;
          } else if (fd.var_slot >= 0) {
            const int32_t ul = A.uniform_len[fd.var_slot];
            if (ul > 0) shape_bad = 1;                                        // a null row has no values: not uniform
            else if (ul < 0) for (int l = 0; l < fd.n_levels; ++l) A.cnt[(size_t)(fd.cnt_slot + l) * A.n + row] = 0;
          }
        }
      }
      if (lane < lim) {
        const uint32_t f = f0 + lane;
        reinterpret_cast<uint32_t*>(A.bitmaps + (size_t)f * A.nb_stride)[blockIdx.x] = my_word;
        const uint32_t nulls = __popc(act_mask & ~my_word);
        if (nulls) atomicAdd(&A.null_counts[f], (unsigned long long)nulls);
      }
    }
  }
  if (bad) atomicOr(A.flags, TF_FALLBACK);
  if (shape_bad) atomicOr(A.flags, TF_SHAPE | TF_FALLBACK);
}

// offsets of a uniform column: offs[i] = i * L  (i = 0..n)
__global__ void uniform_offsets_kernel(int32_t* const* __restrict__ offs, uint32_t stride, const int32_t* __restrict__ uniform_len, uint32_t n_var,
                                       uint32_t n) {
  const uint32_t v = blockIdx.y;
  if (v >= n_var) return;
  const int32_t L = uniform_len[v];
  if (L < 0) return;
  int32_t* o = offs[v * stride];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) o[i] = (int32_t)(i * (uint32_t)L);
}
