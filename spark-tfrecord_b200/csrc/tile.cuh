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
// Mapping: one CTA = one K1 chunk (the records that START in a 32..96 KiB window, contiguous in memory).
//   1. one thread arms an mbarrier and issues cp.async.bulk (TMA bulk copy, SASS UBLKCP) of the window
//      into shared memory; the other threads stage the CRC tables meanwhile.
//   2. thread 0 walks the record chain inside the tile (29-cycle LDS hops instead of ~600 ns DRAM hops).
//   3. thread i parses record i from shared memory: serial slicing-by-8 CRC-32C, strict wire parse,
//      "expected next field" key match with a hash fallback, coercions.
//   4. lane = row, so each column store of a warp covers 32 consecutive rows: coalesced by construction
//      (no shared-memory transpose needed).  Variable-width columns either write element counts + source
//      offsets (then scan + decode_pass2_kernel finish them), or -- when the decoder has learned that a
//      column has a uniform shape (FloatList[8], 16-byte BytesList ...) -- write the values directly at
//      row * L in the same pass and only verify the shape ("uniform-shape speculation").
#pragma once
#include "common.cuh"
#include "decode.cuh"
#include "frame.cuh"

#define TILE_SPILL 12288u       // bytes a record may extend past the end of its chunk and still fit the tile

struct TileArgs {
  const uint8_t* data;
  uint32_t nbytes;
  const ChunkInfo* chunks;
  const uint32_t* chunk_base;   // exclusive prefix of chunk counts = first row of each chunk
  uint32_t n_chunks, chunk_bytes;
  uint32_t n;                   // rows in the batch (stride of the scratch arrays)
  uint32_t verify;
  DevSchema sch;
  const CrcTables* tabs;
  uint8_t* valid8;              // [nf][n]
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

// ---- shared-memory byte helpers -----------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const uint8_t* p) {   // 4 bytes at any alignment
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
  uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t lo = w[0];
  return sh ? __funnelshift_r(lo, w[1], sh) : lo;
}
// varint32 (lengths): accepts 1..5 bytes, minimal or not; false if it runs past `end` or is longer
__device__ __forceinline__ bool s_len(const uint8_t*& p, const uint8_t* end, uint32_t& v) {
  if (p >= end) return false;
  uint32_t b = *p++;
  if (b < 0x80) { v = b; return true; }
  uint32_t r = b & 0x7f;
#pragma unroll 1
  for (int sh = 7; sh < 35; sh += 7) {
    if (p >= end) return false;
    b = *p++;
    r |= (b & 0x7f) << sh;
    if (b < 0x80) { v = r; return (int32_t)r >= 0; }
  }
  return false;   // >5 bytes: leave it to the general path
}

// serial slicing-by-8 CRC-32C over shared memory; s8 = 8 tables of 256 words
__device__ __forceinline__ uint32_t crc_serial_s8(const uint32_t* s8, const uint8_t* p, uint32_t n) {
  uint32_t c = 0xFFFFFFFFu;
  while (n && (reinterpret_cast<uintptr_t>(p) & 3)) { c = (c >> 8) ^ s8[(c ^ *p++) & 0xff]; --n; }
  const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
  uint32_t nw = n >> 3;
#pragma unroll 2
  for (uint32_t i = 0; i < nw; ++i) {
    uint32_t a = w[2 * i] ^ c, b = w[2 * i + 1];
    c = s8[7 * 256 + (a & 0xff)] ^ s8[6 * 256 + ((a >> 8) & 0xff)] ^ s8[5 * 256 + ((a >> 16) & 0xff)] ^ s8[4 * 256 + (a >> 24)] ^
        s8[3 * 256 + (b & 0xff)] ^ s8[2 * 256 + ((b >> 8) & 0xff)] ^ s8[1 * 256 + ((b >> 16) & 0xff)] ^ s8[(b >> 24)];
  }
  p += (size_t)nw * 8; n &= 7;
  while (n--) c = (c >> 8) ^ s8[(c ^ *p++) & 0xff];
  return ~c;
}

// per-thread result of one var-width cell: written in count mode
__device__ __forceinline__ void tile_write_count(const TileArgs& A, const DevField& fd, uint32_t row, uint32_t cnt, uint32_t src_off) {
  A.cnt[(size_t)fd.cnt_slot * A.n + row] = cnt;
  A.src[(size_t)fd.var_slot * A.n + row] = src_off;
  A.cflag[(size_t)fd.var_slot * A.n + row] = CF_CANON;
}

// shared memory layout (dynamic): [0,8) mbarrier | [16, 16+8*1024) CRC tables | rec_off[...] | tile bytes (16-B aligned)
__host__ __device__ inline uint32_t tile_smem_bytes(uint32_t threads, uint32_t chunk_bytes) {
  return 16 + 8192 + ((threads + 1) * 4 + 15) / 16 * 16 + chunk_bytes + TILE_SPILL + 32 + 16;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) decode_tile_kernel(TileArgs A) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* s8 = reinterpret_cast<uint32_t*>(smem_raw + 16);                 // 8 KiB
  uint32_t* rec_offs = reinterpret_cast<uint32_t*>(smem_raw + 16 + 8192);    // THREADS+1 entries per pass
  uint8_t* tile = smem_raw + 16 + 8192 + ((THREADS + 1) * 4 + 15) / 16 * 16;
  const uint32_t tile_cap = A.chunk_bytes + TILE_SPILL + 32;
  const uint32_t nf = (uint32_t)A.sch.n_fields;

  const uint32_t k = blockIdx.x;
  ChunkInfo ci = A.chunks[k];
  if (ci.first == 0xffffffffu || ci.count == 0) return;          // uniform for the CTA
  const uint32_t row0 = A.chunk_base[k];
  // records of this chunk live in [first, end); stop != LEFT means the stream ended/failed in this chunk, the
  // complete records still are [first, end)
  const uint32_t g0 = ci.first & ~15u;
  const uint32_t span = ci.end - g0;
  if (span > tile_cap) {                                          // a record too large for the tile: general path
    if (threadIdx.x == 0) atomicOr(A.flags, TF_FALLBACK);
    return;
  }
  const uint32_t copy_bytes = (span + 15u) & ~15u;                // may read < 16 bytes past `end` (input buffers are padded)
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, copy_bytes);
    uint32_t done = 0;
    while (done < copy_bytes) {                                   // split so each bulk copy stays modest
      uint32_t part = min(copy_bytes - done, 32768u);
      bulk_g2s(tile + done, A.data + g0 + done, part, bar);
      done += part;
    }
  }
  // meanwhile: slicing-by-8 tables into shared memory
  {
    const uint32_t* g = reinterpret_cast<const uint32_t*>(A.tabs->s8);
    for (uint32_t i = threadIdx.x; i < 2048; i += THREADS) s8[i] = g[i];
  }
  mbar_wait(bar, 0);
  __syncthreads();

  uint32_t q = ci.first;     // chain position (thread 0)
  for (uint32_t r0 = 0; r0 < ci.count; r0 += THREADS) {
    const uint32_t cnt_pass = min((uint32_t)THREADS, ci.count - r0);
    if (threadIdx.x == 0) {
      for (uint32_t i = 0; i < cnt_pass; ++i) {
        rec_offs[i] = q;
        q += 16 + s_u32(tile + (q - g0));                         // headers were validated by K1
      }
      rec_offs[cnt_pass] = q;
    }
    __syncthreads();
    if (threadIdx.x < cnt_pass) {
      const uint32_t row = row0 + r0 + threadIdx.x;
      const uint32_t off = rec_offs[threadIdx.x];
      const uint32_t len = rec_offs[threadIdx.x + 1] - off - 16;
      const uint8_t* payload = tile + (off - g0) + 12;
      const uint8_t* end = payload + len;
      bool bad = false;
      uint32_t shape_bad = 0;
      if (A.verify) {
        uint32_t crc = crc_serial_s8(s8, payload, len);
        if (crc_mask(crc) != s_u32(end)) bad = true;
      }
      unsigned long long seen_lo = 0, seen_hi = 0;
      const uint8_t* p = payload;
      uint32_t L = 0;
      // Example { features = 1 }: exactly one field spanning the payload
      if (bad || len < 2 || *p != 0x0A) bad = true;
      else { ++p; if (!s_len(p, end, L) || p + L != end) bad = true; }
      uint32_t next_f = 0;
      while (!bad && p < end) {
        // ---- map entry ----
        uint32_t elen, klen, vlen;
        if (*p != 0x0A) { bad = true; break; }
        ++p;
        if (!s_len(p, end, elen) || (uint32_t)(end - p) < elen) { bad = true; break; }
        const uint8_t* eend = p + elen;
        if (p >= eend || *p != 0x0A) { bad = true; break; }
        ++p;
        if (!s_len(p, eend, klen) || (uint32_t)(eend - p) < klen) { bad = true; break; }
        const uint8_t* key = p;
        p += klen;
        if (p >= eend || *p != 0x12) { bad = true; break; }
        ++p;
        if (!s_len(p, eend, vlen) || p + vlen != eend) { bad = true; break; }
        // ---- which schema field? entries normally arrive in schema order ----
        int f = -1;
        if (next_f < nf) {
          const DevField& nd = A.sch.fields[next_f];
          if (nd.name_len == klen) {
            const uint8_t* nm = A.sch.names + nd.name_off;
            bool eq = true;
            for (uint32_t i = 0; i < klen; ++i) if (nm[i] != key[i]) { eq = false; break; }
            if (eq) f = (int)next_f;
          }
        }
        if (f < 0) f = schema_lookup(A.sch, key, klen);
        {
          uint32_t hi = 0;
          for (uint32_t i = 0; i < klen; ++i) hi |= key[i];
          if (hi >= 0x80 && !utf8_valid(key, klen)) { bad = true; break; }     // malformed key: the general path reports it
        }
        if (f >= 0 && A.sch.fields[f].elem_type == TFR_T_NULL) f = -1;           // NullType: always null, value only validated
        if (f >= 0) {
          unsigned long long bit = 1ull << (f & 63);
          unsigned long long& s = (f < 64) ? seen_lo : seen_hi;
          if (s & bit) { bad = true; break; }                                   // duplicate key: last-wins semantics -> general path
          s |= bit;
          next_f = (uint32_t)f + 1;
        }
        // ---- Feature: exactly one oneof member spanning the value ----
        if (vlen == 0) { if (f >= 0) bad = true; p = eend; continue; }         // kind not set: an error if the schema wants it
        uint32_t kt = *p++, llen;
        uint32_t kind = kt == 0x0A ? K_BYTES : kt == 0x12 ? K_FLOAT : kt == 0x1A ? K_INT64 : K_NONE;
        if (kind == K_NONE || !s_len(p, eend, llen) || p + llen != eend) { bad = true; break; }
        const DevField* fd = f >= 0 ? &A.sch.fields[f] : nullptr;
        if (fd && ((uint32_t)fd->kind != kind || fd->depth > 1)) { bad = true; break; }   // kind mismatch / nesting: error path
        if (kind == K_BYTES) {
          // BytesList: { 0A blen bytes }*
          uint32_t n = 0, first_off = 0, first_len = 0, total = 0;
          const uint8_t* body = p;
          const uint8_t* first_data = p;
          const bool is_str = fd && fd->elem_type == TFR_T_STRING;
          while (p < eend) {
            uint32_t bl;
            if (*p != 0x0A) { bad = true; break; }
            ++p;
            const uint8_t* lp = p;
            if (!s_len(p, eend, bl) || (uint32_t)(eend - p) < bl) { bad = true; break; }
            if (is_str) {
              // StringType goes through Java's UTF-8 decode/re-encode: the identity for well-formed input; malformed
              // input needs the U+FFFD transcode, which only the general path implements
              uint32_t acc = 0;
              for (uint32_t i = 0; i < bl; ++i) acc |= p[i];
              if (acc >= 0x80 && !utf8_valid(p, bl)) { bad = true; break; }
            }
            if (n == 0) { first_off = (uint32_t)(lp - tile) + g0; first_len = bl; first_data = p; }
            ++n; total += bl;
            p += bl;
          }
          if (bad) break;
          if (fd) {
            if (fd->depth == 0) {
              if (n == 0) { bad = true; break; }                                // .head of an empty list: error path
              int32_t ul = A.uniform_len[fd->var_slot];
              if (ul >= 0) {
                if ((uint32_t)ul != first_len) shape_bad = 1;
                else {
                  uint8_t* dst = reinterpret_cast<uint8_t*>(A.var_values[fd->var_slot]) + (size_t)row * (uint32_t)ul;
                  const uint8_t* sp = first_data;
                  if ((ul & 15) == 0) {
                    for (uint32_t i = 0; i < (uint32_t)ul; i += 16) {
                      uint4 v; v.x = s_u32(sp + i); v.y = s_u32(sp + i + 4); v.z = s_u32(sp + i + 8); v.w = s_u32(sp + i + 12);
                      *reinterpret_cast<uint4*>(dst + i) = v;
                    }
                  } else for (uint32_t i = 0; i < (uint32_t)ul; ++i) dst[i] = sp[i];
                }
              } else tile_write_count(A, *fd, row, first_len, first_off);
            } else {
              // ArrayType(String/Binary): two offset levels -> always count mode (the decoder never speculates on these)
              A.cnt[(size_t)fd->cnt_slot * A.n + row] = n;
              A.cnt[(size_t)(fd->cnt_slot + 1) * A.n + row] = total;
              A.src[(size_t)fd->var_slot * A.n + row] = (uint32_t)(body - tile) + g0;
              A.cflag[(size_t)fd->var_slot * A.n + row] = CF_CANON;
            }
          }
        } else {
          // Int64List / FloatList: empty, or one packed field spanning the list
          uint32_t plen = 0;
          const uint8_t* pk = p;
          if (llen != 0) {
            if (*p != 0x0A) { bad = true; break; }
            ++p;
            if (!s_len(p, eend, plen) || p + plen != eend) { bad = true; break; }
            pk = p;
          }
          uint32_t n;
          if (kind == K_FLOAT) {
            if (plen & 3) { bad = true; break; }
            n = plen >> 2;
          } else {
            // varints: count terminators, every run <= 10 bytes, last byte terminates
            n = 0; uint32_t run = 0;
            for (uint32_t i = 0; i < plen; ++i) {
              if (pk[i] & 0x80) { if (++run >= 10) { bad = true; break; } } else { ++n; run = 0; }
            }
            if (bad || run) { bad = true; break; }
          }
          if (fd) {
            if (fd->depth == 0) {
              if (n == 0) { bad = true; break; }
              void* vp = A.fix_values[fd->fix_slot];
              if (kind == K_FLOAT) {
                uint32_t bits = s_u32(pk);
                if (fd->elem_type == TFR_T_FLOAT32) reinterpret_cast<uint32_t*>(vp)[row] = bits;
                else reinterpret_cast<double*>(vp)[row] = (double)__uint_as_float(bits);
              } else {
                uint64_t v = 0; uint32_t sh = 0;
                for (uint32_t i = 0; i < 10; ++i) { uint32_t b = pk[i]; v |= (uint64_t)(b & 0x7f) << sh; sh += 7; if (b < 0x80) break; }
                if (fd->elem_type == TFR_T_INT64) reinterpret_cast<int64_t*>(vp)[row] = (int64_t)v;
                else reinterpret_cast<int32_t*>(vp)[row] = (int32_t)(uint32_t)v;
              }
            } else {
              int32_t ul = A.uniform_len[fd->var_slot];
              if (ul >= 0) {
                if ((uint32_t)ul != n) shape_bad = 1;
                else if (kind == K_FLOAT) {
                  if (fd->elem_type == TFR_T_FLOAT32) {
                    uint32_t* dst = reinterpret_cast<uint32_t*>(A.var_values[fd->var_slot]) + (size_t)row * n;
                    if ((n & 3) == 0) {
                      for (uint32_t i = 0; i < n; i += 4) {
                        uint4 v; v.x = s_u32(pk + 4 * i); v.y = s_u32(pk + 4 * i + 4); v.z = s_u32(pk + 4 * i + 8); v.w = s_u32(pk + 4 * i + 12);
                        *reinterpret_cast<uint4*>(dst + i) = v;
                      }
                    } else for (uint32_t i = 0; i < n; ++i) dst[i] = s_u32(pk + 4 * i);
                  } else {
                    double* dst = reinterpret_cast<double*>(A.var_values[fd->var_slot]) + (size_t)row * n;
                    for (uint32_t i = 0; i < n; ++i) dst[i] = (double)__uint_as_float(s_u32(pk + 4 * i));
                  }
                } else {
                  const uint8_t* vpz = pk;
                  for (uint32_t e = 0; e < n; ++e) {
                    uint64_t v = 0; uint32_t sh = 0;
                    for (;;) { uint32_t b = *vpz++; v |= (uint64_t)(b & 0x7f) << sh; sh += 7; if (b < 0x80) break; }
                    if (fd->elem_type == TFR_T_INT64) reinterpret_cast<int64_t*>(A.var_values[fd->var_slot])[(size_t)row * n + e] = (int64_t)v;
                    else reinterpret_cast<int32_t*>(A.var_values[fd->var_slot])[(size_t)row * n + e] = (int32_t)(uint32_t)v;
                  }
                }
              } else tile_write_count(A, *fd, row, n, (uint32_t)(pk - tile) + g0);
            }
          }
          p = eend;
        }
        p = eend;
      }
      // ---- absent fields: null (or an error for non-nullable ones), validity bytes ----
      if (!bad) {
        for (uint32_t f = 0; f < nf; ++f) {
          const DevField& fd = A.sch.fields[f];
          bool present = f < 64 ? (seen_lo >> f) & 1 : (seen_hi >> (f - 64)) & 1;
          A.valid8[(size_t)f * A.n + row] = present ? 1 : 0;
          if (!present) {
            if (!fd.nullable && fd.elem_type != TFR_T_NULL) { bad = true; }
            if (fd.elem_type == TFR_T_NULL) continue;
            if (fd.fix_slot >= 0) {
              void* vp = A.fix_values[fd.fix_slot];
              if (fd.width == 8) reinterpret_cast<uint64_t*>(vp)[row] = 0; else reinterpret_cast<uint32_t*>(vp)[row] = 0;
            } else if (fd.var_slot >= 0) {
              int32_t ul = A.uniform_len[fd.var_slot];
              if (ul > 0) shape_bad = 1;                                        // a null row has no values: shape is not uniform
              else if (ul < 0) for (int l = 0; l < fd.n_levels; ++l) A.cnt[(size_t)(fd.cnt_slot + l) * A.n + row] = 0;
            }
          }
        }
      }
      if (bad) atomicOr(A.flags, TF_FALLBACK);
      if (shape_bad) atomicOr(A.flags, TF_SHAPE | TF_FALLBACK);
    }
    __syncthreads();
  }
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
