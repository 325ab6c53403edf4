/*
 * tfr_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this.  The product (libtfrgpu.so) never links, loads or calls it.
 *
 * What it is: a plain-C restatement of the reference's CPU algorithm for the TFRecord
 * decode/encode hot path of linkedin/spark-tfrecord @ 5bc46ee, keeping its algorithmic shape
 * (one record at a time; table CRC; full object-tree parse; per-schema-field name lookup;
 * row-at-a-time materialisation), so it doubles as the timed CPU baseline ("kind": "port").
 *
 * PARITY PINNING: the reference cannot run here (no JVM) and its own tests hold no golden bytes
 *  -> byte-level parity with the reference is "parity unpinned".  The oracle is
 * pinned instead against (a) RFC 3720 CRC-32C check values and the TFRecord framing spec,
 * (b) an independent protobuf implementation (google.protobuf/upb, tests/test_oracle_vs_upb.py),
 * (c) every literal case of the reference's scalatest suites restated in tests/.
 *
 * Reference shorthand: M/ = src/main/scala/com/linkedin/spark/datasources/tfrecord/
 * Third-party (unvendored, restated from their published behaviour):
 *   org.tensorflow:tensorflow-hadoop:1.15.0  (TFRecordReader / TFRecordWriter / Crc32C)
 *   com.google.protobuf:protobuf-java 3.x    (CodedInputStream / generated Example parsers)
 */
#include "../include/tfrgpu.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ======================================================================================
 * CRC-32C, slicing-by-8 (tensorflow-hadoop Crc32C = Hadoop PureJavaCrc32C: 8 tables of 256)
 * ====================================================================================== */
static uint32_t T8[8][256];
static int crc_init_done = 0;
static void crc_init(void) {
  if (crc_init_done) return;
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
    T8[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; i++)
    for (int t = 1; t < 8; t++) T8[t][i] = (T8[t - 1][i] >> 8) ^ T8[0][T8[t - 1][i] & 0xff];
  crc_init_done = 1;
}
uint32_t tfr_oracle_crc32c(const uint8_t* p, size_t n) {
  crc_init();
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint32_t lo = c ^ ((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24);
    c = T8[7][lo & 0xff] ^ T8[6][(lo >> 8) & 0xff] ^ T8[5][(lo >> 16) & 0xff] ^ T8[4][lo >> 24] ^
        T8[3][p[4]] ^ T8[2][p[5]] ^ T8[1][p[6]] ^ T8[0][p[7]];
    p += 8; n -= 8;
  }
  while (n--) c = (c >> 8) ^ T8[0][(c ^ *p++) & 0xff];
  return c ^ 0xFFFFFFFFu;
}
/* Crc32C.getMaskedValue(): ((v >>> 15) | (v << 17)) + 0xa282ead8 */
uint32_t tfr_oracle_masked_crc32c(const uint8_t* p, size_t n) {
  uint32_t v = tfr_oracle_crc32c(p, n);
  return ((v >> 15) | (v << 17)) + 0xa282ead8u;
}

/* ======================================================================================
 * small growable buffers
 * ====================================================================================== */
typedef struct { uint8_t* p; size_t n, cap; } buf_t;
static void buf_reserve(buf_t* b, size_t extra) {
  if (b->n + extra <= b->cap) return;
  size_t nc = b->cap ? b->cap * 2 : 256;
  while (nc < b->n + extra) nc *= 2;
  b->p = (uint8_t*)realloc(b->p, nc);
  b->cap = nc;
}
static void buf_put(buf_t* b, const void* src, size_t n) {
  buf_reserve(b, n);
  if (n) memcpy(b->p + b->n, src, n);
  b->n += n;
}
static void buf_put_i32(buf_t* b, int32_t v) { buf_put(b, &v, 4); }
static void buf_put_u8(buf_t* b, uint8_t v) { buf_put(b, &v, 1); }
static void buf_free(buf_t* b) { free(b->p); b->p = NULL; b->n = b->cap = 0; }

/* ======================================================================================
 * protobuf object model (org.tensorflow.example.*), parsed the way protobuf-java does
 * ====================================================================================== */
enum { K_NONE = 0, K_BYTES = 1, K_FLOAT = 2, K_INT64 = 3 };   /* Feature.KindCase numbers */

typedef struct { uint8_t* p; uint32_t n; } bytes_t;            /* ByteString (copied)     */
typedef struct {
  int kind;
  /* BytesList.value / FloatList.value / Int64List.value */
  bytes_t* bytes; float* floats; int64_t* ints; /* floats kept as raw bit patterns via memcpy */
  size_t n, cap;
} feature_t;

typedef struct { feature_t* f; size_t n, cap; } featurelist_t;  /* FeatureList.feature */

typedef struct {
  char* key; uint32_t key_len; uint32_t hash;
  feature_t feat;          /* value when this is a Features map */
  featurelist_t flist;     /* value when this is a FeatureLists map */
} mapent_t;
/* LinkedHashMap<String, V>: insertion-ordered entries + open-addressing index */
typedef struct { mapent_t* e; size_t n, cap; int32_t* idx; size_t idx_cap; int is_flist; } pbmap_t;

typedef struct { const uint8_t* p; const uint8_t* end; int depth; } cis_t; /* CodedInputStream */

static void feature_clear(feature_t* f) {
  if (f->kind == K_BYTES) for (size_t i = 0; i < f->n; i++) free(f->bytes[i].p);
  free(f->bytes); free(f->floats); free(f->ints);
  memset(f, 0, sizeof *f);
}
static void flist_clear(featurelist_t* l) {
  for (size_t i = 0; i < l->n; i++) feature_clear(&l->f[i]);
  free(l->f); memset(l, 0, sizeof *l);
}
static void map_clear(pbmap_t* m) {
  for (size_t i = 0; i < m->n; i++) {
    free(m->e[i].key);
    feature_clear(&m->e[i].feat);
    flist_clear(&m->e[i].flist);
  }
  free(m->e); free(m->idx);
  int fl = m->is_flist; memset(m, 0, sizeof *m); m->is_flist = fl;
}
static uint32_t str_hash(const char* s, uint32_t n) {   /* String.hashCode shape: h*31 + c */
  uint32_t h = 0; for (uint32_t i = 0; i < n; i++) h = h * 31u + (uint8_t)s[i]; return h;
}
static void map_reindex(pbmap_t* m) {
  size_t nc = 16; while (nc < 2 * (m->n + 1)) nc *= 2;
  free(m->idx); m->idx = (int32_t*)malloc(nc * sizeof(int32_t)); m->idx_cap = nc;
  for (size_t i = 0; i < nc; i++) m->idx[i] = -1;
  for (size_t i = 0; i < m->n; i++) {
    size_t h = m->e[i].hash & (nc - 1);
    while (m->idx[h] >= 0) h = (h + 1) & (nc - 1);
    m->idx[h] = (int32_t)i;
  }
}
static mapent_t* map_get(const pbmap_t* m, const char* key, uint32_t n) {
  if (!m->idx_cap) return NULL;
  uint32_t hv = str_hash(key, n);
  size_t h = hv & (m->idx_cap - 1);
  while (m->idx[h] >= 0) {
    mapent_t* e = &m->e[m->idx[h]];
    if (e->hash == hv && e->key_len == n && (n == 0 || memcmp(e->key, key, n) == 0)) return e;
    h = (h + 1) & (m->idx_cap - 1);
  }
  return NULL;
}
/* Map.put: replace the value of an existing key (keeping its position) or append */
static mapent_t* map_put_slot(pbmap_t* m, const char* key, uint32_t n) {
  mapent_t* e = map_get(m, key, n);
  if (e) { feature_clear(&e->feat); flist_clear(&e->flist); return e; }
  if (m->n == m->cap) { m->cap = m->cap ? m->cap * 2 : 16; m->e = (mapent_t*)realloc(m->e, m->cap * sizeof(mapent_t)); }
  e = &m->e[m->n++];
  memset(e, 0, sizeof *e);
  e->key = (char*)malloc(n ? n : 1); if (n) memcpy(e->key, key, n);
  e->key_len = n; e->hash = str_hash(key, n);
  if (2 * (m->n + 1) > m->idx_cap) map_reindex(m);
  else { size_t h = e->hash & (m->idx_cap - 1); while (m->idx[h] >= 0) h = (h + 1) & (m->idx_cap - 1); m->idx[h] = (int32_t)(m->n - 1); }
  return e;
}

/* ---- CodedInputStream primitives (protobuf-java CodedInputStream.ArrayDecoder) ---------- */
#define PB_OK 0
#define PB_ERR 1
/* readRawVarint64: at most 10 bytes, low 64 bits kept; 10 continuation bytes -> malformedVarint */
static int rd_varint64(cis_t* s, uint64_t* out) {
  uint64_t v = 0;
  for (int i = 0; i < 10; i++) {
    if (s->p >= s->end) return PB_ERR;            /* truncatedMessage */
    uint8_t b = *s->p++;
    if (i < 9) v |= (uint64_t)(b & 0x7f) << (7 * i);
    else v |= (uint64_t)(b & 0x01) << 63;
    if (!(b & 0x80)) { *out = v; return PB_OK; }
  }
  return PB_ERR;                                  /* malformedVarint */
}
/* readRawVarint32: same bytes consumed, low 32 bits kept */
static int rd_varint32(cis_t* s, uint32_t* out) {
  uint64_t v; if (rd_varint64(s, &v)) return PB_ERR; *out = (uint32_t)v; return PB_OK;
}
/* readTag: 0 at end of the current limit; field number 0 -> invalidTag */
static int rd_tag(cis_t* s, uint32_t* tag) {
  if (s->p >= s->end) { *tag = 0; return PB_OK; }
  if (rd_varint32(s, tag)) return PB_ERR;
  if ((*tag >> 3) == 0) return PB_ERR;
  return PB_OK;
}
/* length prefix + pushLimit: negative -> negativeSize, beyond limit -> truncatedMessage */
static int rd_len(cis_t* s, uint32_t* len) {
  if (rd_varint32(s, len)) return PB_ERR;
  if ((int32_t)*len < 0) return PB_ERR;
  if ((size_t)(s->end - s->p) < *len) return PB_ERR;
  return PB_OK;
}
#define PB_RECURSION_LIMIT 100
/* UnknownFieldSet.Builder.mergeFieldFrom: returns 1 when the tag was END_GROUP (caller stops) */
static int skip_field(cis_t* s, uint32_t tag, int* end_group) {
  *end_group = 0;
  switch (tag & 7) {
    case 0: { uint64_t v; return rd_varint64(s, &v); }
    case 1: if (s->end - s->p < 8) return PB_ERR; s->p += 8; return PB_OK;
    case 2: { uint32_t l; if (rd_len(s, &l)) return PB_ERR; s->p += l; return PB_OK; }
    case 3: {  /* START_GROUP: nested unknown fields until the matching END_GROUP */
      if (++s->depth > PB_RECURSION_LIMIT) return PB_ERR;
      for (;;) {
        uint32_t t; if (rd_tag(s, &t)) return PB_ERR;
        if (t == 0) return PB_ERR;                 /* limit hit inside the group */
        int eg; if (skip_field(s, t, &eg)) return PB_ERR;
        if (eg) { if ((t >> 3) != (tag >> 3)) return PB_ERR; break; }  /* checkLastTagWas */
      }
      s->depth--;
      return PB_OK;
    }
    case 4: *end_group = 1; return PB_OK;
    case 5: if (s->end - s->p < 4) return PB_ERR; s->p += 4; return PB_OK;
    default: return PB_ERR;                        /* invalidWireType */
  }
}

/* isValidUtf8 (protobuf Utf8.isValidUtf8: well-formed per Unicode, no surrogates, no overlongs) */
static int utf8_valid(const uint8_t* p, size_t n) {
  size_t i = 0;
  while (i < n) {
    uint8_t b = p[i];
    if (b < 0x80) { i++; continue; }
    if (b < 0xC2) return 0;
    if (b < 0xE0) { if (i + 1 >= n || (p[i + 1] & 0xC0) != 0x80) return 0; i += 2; continue; }
    if (b < 0xF0) {
      if (i + 2 >= n) return 0;
      uint8_t b2 = p[i + 1], b3 = p[i + 2];
      if ((b2 & 0xC0) != 0x80 || (b3 & 0xC0) != 0x80) return 0;
      if (b == 0xE0 && b2 < 0xA0) return 0;
      if (b == 0xED && b2 >= 0xA0) return 0;
      i += 3; continue;
    }
    if (b > 0xF4) return 0;
    if (i + 3 >= n) return 0;
    uint8_t b2 = p[i + 1], b3 = p[i + 2], b4 = p[i + 3];
    if ((b2 & 0xC0) != 0x80 || (b3 & 0xC0) != 0x80 || (b4 & 0xC0) != 0x80) return 0;
    if (b == 0xF0 && b2 < 0x90) return 0;
    if (b == 0xF4 && b2 >= 0x90) return 0;
    i += 4;
  }
  return 1;
}

static void feat_grow(feature_t* f, size_t extra) {
  if (f->n + extra <= f->cap) return;
  size_t nc = f->cap ? f->cap * 2 : 4; while (nc < f->n + extra) nc *= 2;
  if (f->kind == K_BYTES) f->bytes = (bytes_t*)realloc(f->bytes, nc * sizeof(bytes_t));
  else if (f->kind == K_FLOAT) f->floats = (float*)realloc(f->floats, nc * sizeof(float));
  else f->ints = (int64_t*)realloc(f->ints, nc * sizeof(int64_t));
  f->cap = nc;
}
/* one occurrence of a oneof kind field inside Feature: switching kind discards the previous
 * value, the same kind again merges (repeated fields concatenate) */
static int parse_list_into(cis_t* s, feature_t* f, int kind) {
  uint32_t len; if (rd_len(s, &len)) return PB_ERR;
  if (++s->depth > PB_RECURSION_LIMIT) return PB_ERR;
  cis_t sub = { s->p, s->p + len, s->depth };
  s->p += len;
  if (f->kind != kind) { feature_clear(f); f->kind = kind; }
  for (;;) {
    uint32_t tag; if (rd_tag(&sub, &tag)) return PB_ERR;
    if (tag == 0) break;
    if (kind == K_BYTES && tag == 0x0A) {
      uint32_t l; if (rd_len(&sub, &l)) return PB_ERR;
      feat_grow(f, 1);
      f->bytes[f->n].p = (uint8_t*)malloc(l ? l : 1); memcpy(f->bytes[f->n].p, sub.p, l);
      f->bytes[f->n].n = l; f->n++; sub.p += l;
    } else if (kind == K_FLOAT && tag == 0x0D) {          /* unpacked fixed32 */
      if (sub.end - sub.p < 4) return PB_ERR;
      feat_grow(f, 1); memcpy(&f->floats[f->n++], sub.p, 4); sub.p += 4;
    } else if (kind == K_FLOAT && tag == 0x0A) {          /* packed */
      uint32_t l; if (rd_len(&sub, &l)) return PB_ERR;
      /* while (getBytesUntilLimit() > 0) readFloat(): a ragged tail is truncatedMessage */
      if (l % 4) return PB_ERR;
      feat_grow(f, l / 4); memcpy(&f->floats[f->n], sub.p, l); f->n += l / 4; sub.p += l;
    } else if (kind == K_INT64 && tag == 0x08) {          /* unpacked varint */
      uint64_t v; if (rd_varint64(&sub, &v)) return PB_ERR;
      feat_grow(f, 1); f->ints[f->n++] = (int64_t)v;
    } else if (kind == K_INT64 && tag == 0x0A) {          /* packed */
      uint32_t l; if (rd_len(&sub, &l)) return PB_ERR;
      cis_t pk = { sub.p, sub.p + l, sub.depth }; sub.p += l;
      while (pk.p < pk.end) {
        uint64_t v; if (rd_varint64(&pk, &v)) return PB_ERR;
        feat_grow(f, 1); f->ints[f->n++] = (int64_t)v;
      }
    } else {
      int eg; if (skip_field(&sub, tag, &eg)) return PB_ERR;
      if (eg) return PB_ERR;                               /* checkLastTagWas(0) fails */
    }
  }
  s->depth--;
  return PB_OK;
}
/* Feature message body (merging into *f, as Feature.Builder.mergeFrom does) */
static int parse_feature_body(cis_t* sub, feature_t* f) {
  for (;;) {
    uint32_t tag; if (rd_tag(sub, &tag)) return PB_ERR;
    if (tag == 0) return PB_OK;
    if (tag == 0x0A) { if (parse_list_into(sub, f, K_BYTES)) return PB_ERR; }
    else if (tag == 0x12) { if (parse_list_into(sub, f, K_FLOAT)) return PB_ERR; }
    else if (tag == 0x1A) { if (parse_list_into(sub, f, K_INT64)) return PB_ERR; }
    else { int eg; if (skip_field(sub, tag, &eg)) return PB_ERR; if (eg) return PB_ERR; }
  }
}
static int parse_feature(cis_t* s, feature_t* f) {
  uint32_t len; if (rd_len(s, &len)) return PB_ERR;
  if (++s->depth > PB_RECURSION_LIMIT) return PB_ERR;
  cis_t sub = { s->p, s->p + len, s->depth }; s->p += len;
  if (parse_feature_body(&sub, f)) return PB_ERR;
  s->depth--;
  return PB_OK;
}
/* FeatureList message: repeated Feature feature = 1 (merge = append) */
static int parse_featurelist(cis_t* s, featurelist_t* l) {
  uint32_t len; if (rd_len(s, &len)) return PB_ERR;
  if (++s->depth > PB_RECURSION_LIMIT) return PB_ERR;
  cis_t sub = { s->p, s->p + len, s->depth }; s->p += len;
  for (;;) {
    uint32_t tag; if (rd_tag(&sub, &tag)) return PB_ERR;
    if (tag == 0) break;
    if (tag == 0x0A) {
      if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 8; l->f = (feature_t*)realloc(l->f, l->cap * sizeof(feature_t)); }
      memset(&l->f[l->n], 0, sizeof(feature_t));
      l->n++;
      if (parse_feature(&sub, &l->f[l->n - 1])) return PB_ERR;
    } else { int eg; if (skip_field(&sub, tag, &eg)) return PB_ERR; if (eg) return PB_ERR; }
  }
  s->depth--;
  return PB_OK;
}
/* Features / FeatureLists message: map<string, V> field 1.  Each entry is parsed as a
 * MapEntry message {1: key, 2: value}: last key wins, a repeated value field merges into the
 * entry's current value (MapEntryLite.parseField -> value.toBuilder().mergeFrom), unknown
 * fields are skipped, then map.put(key, value) (last entry wins per key).                  */
static int parse_map_msg(cis_t* s, pbmap_t* m) {
  uint32_t len; if (rd_len(s, &len)) return PB_ERR;
  if (++s->depth > PB_RECURSION_LIMIT) return PB_ERR;
  cis_t sub = { s->p, s->p + len, s->depth }; s->p += len;
  for (;;) {
    uint32_t tag; if (rd_tag(&sub, &tag)) return PB_ERR;
    if (tag == 0) break;
    if (tag != 0x0A) { int eg; if (skip_field(&sub, tag, &eg)) return PB_ERR; if (eg) return PB_ERR; continue; }
    uint32_t elen; if (rd_len(&sub, &elen)) return PB_ERR;
    if (++sub.depth > PB_RECURSION_LIMIT) return PB_ERR;
    cis_t ent = { sub.p, sub.p + elen, sub.depth }; sub.p += elen;
    const uint8_t* key = (const uint8_t*)""; uint32_t klen = 0;
    feature_t fv; memset(&fv, 0, sizeof fv);
    featurelist_t lv; memset(&lv, 0, sizeof lv);
    int rc = PB_OK;
    for (;;) {
      uint32_t t; if (rd_tag(&ent, &t)) { rc = PB_ERR; break; }
      if (t == 0) break;
      if (t == 0x0A) {
        uint32_t l; if (rd_len(&ent, &l)) { rc = PB_ERR; break; }
        /* proto3 string: readStringRequireUtf8 */
        if (!utf8_valid(ent.p, l)) { rc = PB_ERR; break; }
        key = ent.p; klen = l; ent.p += l;
      } else if (t == 0x12) {
        if (m->is_flist ? parse_featurelist(&ent, &lv) : parse_feature(&ent, &fv)) { rc = PB_ERR; break; }
      } else { int eg; if (skip_field(&ent, t, &eg) || eg) { rc = PB_ERR; break; } }
    }
    if (rc) { feature_clear(&fv); flist_clear(&lv); return PB_ERR; }
    sub.depth--;
    mapent_t* e = map_put_slot(m, (const char*)key, klen);
    e->feat = fv; e->flist = lv;
  }
  s->depth--;
  return PB_OK;
}

typedef struct { pbmap_t features; pbmap_t feature_lists; } record_t;  /* Example or SequenceExample */

/* Example.parseFrom(bytes) (M/TFRecordFileReader.scala:73) /
 * SequenceExample.parseFrom(bytes) (:76).  A repeated `features`/`context` field merges. */
static int parse_record(const uint8_t* p, size_t n, int record_type, record_t* r) {
  cis_t s = { p, p + n, 0 };
  for (;;) {
    uint32_t tag; if (rd_tag(&s, &tag)) return PB_ERR;
    if (tag == 0) return PB_OK;
    if (tag == 0x0A) { if (parse_map_msg(&s, &r->features)) return PB_ERR; }
    else if (tag == 0x12 && record_type == TFR_RT_SEQUENCE_EXAMPLE) { if (parse_map_msg(&s, &r->feature_lists)) return PB_ERR; }
    else { int eg; if (skip_field(&s, tag, &eg)) return PB_ERR; if (eg) return PB_ERR; }
  }
}

/* ======================================================================================
 * Java UTF-8 decode + re-encode: ByteString.toStringUtf8 -> UTF8String.fromString
 * (M/TFRecordDeserializer.scala:91,169,215).  Restated from OpenJDK java.lang.String
 * decodeUTF8 (doReplace=true) / sun.nio.cs.UTF_8 malformedN: malformed input becomes
 * U+FFFD (EF BF BD), well-formed input is returned byte-identical.
 * ====================================================================================== */
static int not_cont(uint8_t b) { return (b & 0xc0) != 0x80; }
static void java_utf8_roundtrip(const uint8_t* src, size_t sl, buf_t* out) {
  static const uint8_t REPL[3] = { 0xEF, 0xBF, 0xBD };
  size_t sp = 0;
  while (sp < sl) {
    uint8_t b1 = src[sp];
    if (b1 < 0x80) { buf_put_u8(out, b1); sp++; continue; }
    if ((b1 >> 5) == 0x6 && (b1 & 0x1e) != 0) {             /* 2 bytes: C2..DF */
      if (sp + 1 < sl) {
        uint8_t b2 = src[sp + 1];
        if (not_cont(b2)) { buf_put(out, REPL, 3); sp += 1; }
        else { buf_put(out, src + sp, 2); sp += 2; }
        continue;
      }
      buf_put(out, REPL, 3); break;                           /* truncated tail */
    }
    if ((b1 >> 4) == 0xE) {                                   /* 3 bytes: E0..EF */
      if (sp + 2 < sl) {
        uint8_t b2 = src[sp + 1], b3 = src[sp + 2];
        int mal3 = (b1 == 0xe0 && (b2 & 0xe0) == 0x80) || not_cont(b2) || not_cont(b3);
        if (mal3) {
          buf_put(out, REPL, 3);
          /* malformedN(3): 1 if (b1==E0 && (b2&E0)==80) || notCont(b2) else 2 */
          sp += ((b1 == 0xe0 && (b2 & 0xe0) == 0x80) || not_cont(b2)) ? 1 : 2;
        } else {
          uint32_t c = ((uint32_t)(b1 & 0x0f) << 12) | ((uint32_t)(b2 & 0x3f) << 6) | (b3 & 0x3f);
          if (c >= 0xD800 && c <= 0xDFFF) buf_put(out, REPL, 3);   /* surrogate: one REPL for 3 bytes */
          else buf_put(out, src + sp, 3);
          sp += 3;
        }
        continue;
      }
      if (sp + 1 < sl && ((b1 == 0xe0 && (src[sp + 1] & 0xe0) == 0x80) || not_cont(src[sp + 1]))) {
        buf_put(out, REPL, 3); sp += 1; continue;
      }
      buf_put(out, REPL, 3); break;
    }
    if ((b1 >> 3) == 0x1E) {                                  /* 4 bytes: F0..F7 */
      if (sp + 3 < sl) {
        uint8_t b2 = src[sp + 1], b3 = src[sp + 2], b4 = src[sp + 3];
        uint32_t uc = ((uint32_t)(b1 & 0x07) << 18) | ((uint32_t)(b2 & 0x3f) << 12) | ((uint32_t)(b3 & 0x3f) << 6) | (b4 & 0x3f);
        int mal4 = not_cont(b2) || not_cont(b3) || not_cont(b4);
        if (mal4 || !(uc >= 0x10000 && uc <= 0x10FFFF)) {
          buf_put(out, REPL, 3);
          /* malformedN(4) */
          if (b1 > 0xf4 || (b1 == 0xf0 && (b2 < 0x90 || b2 > 0xbf)) || (b1 == 0xf4 && (b2 & 0xf0) != 0x80) || not_cont(b2)) sp += 1;
          else if (not_cont(b3)) sp += 2;
          else sp += 3;
        } else { buf_put(out, src + sp, 4); sp += 4; }
        continue;
      }
      /* fewer than 4 bytes left */
      {
        uint8_t b2 = sp + 1 < sl ? src[sp + 1] : 0;
        if (b1 > 0xf4 || (sp + 1 < sl && ((b1 == 0xf0 && (b2 < 0x90 || b2 > 0xbf)) || (b1 == 0xf4 && (b2 & 0xf0) != 0x80) || not_cont(b2)))) {
          buf_put(out, REPL, 3); sp += 1; continue;
        }
        if (sp + 2 < sl && not_cont(src[sp + 2])) { buf_put(out, REPL, 3); sp += 2; continue; }
        buf_put(out, REPL, 3); break;
      }
    }
    buf_put(out, REPL, 3); sp += 1;                           /* 80..C1, F8..FF */
  }
}

/* ======================================================================================
 * column builders (Arrow layout == tfr_column with host pointers)
 * ====================================================================================== */
typedef struct {
  int elem_type, depth, n_levels, width;
  buf_t valid;          /* one byte per row while building */
  buf_t off[3];         /* int32 offsets per level         */
  buf_t values;
  int64_t n_rows, null_count;
  uint8_t* bitmap;
} colb_t;

static int type_width(int t) {
  switch (t) { case TFR_T_INT32: case TFR_T_FLOAT32: return 4; case TFR_T_INT64: case TFR_T_FLOAT64: case TFR_T_DECIMAL: return 8;
               case TFR_T_STRING: case TFR_T_BINARY: return 1; default: return 0; }
}
static int is_varlen(int t) { return t == TFR_T_STRING || t == TFR_T_BINARY; }
static void colb_init(colb_t* c, int elem_type, int depth) {
  memset(c, 0, sizeof *c);
  c->elem_type = elem_type; c->depth = depth; c->width = type_width(elem_type);
  c->n_levels = depth + (is_varlen(elem_type) ? 1 : 0);
  for (int l = 0; l < c->n_levels; l++) buf_put_i32(&c->off[l], 0);
}
static int64_t colb_level_count(colb_t* c, int level) { /* children appended so far at `level` */
  if (level == c->n_levels) return (int64_t)(c->values.n / (c->width ? c->width : 1));
  return (int64_t)(c->off[level].n / 4) - 1;
}
static int colb_close(colb_t* c, int level) {  /* push the end offset of one parent at `level` */
  int64_t cnt = colb_level_count(c, level + 1);
  if (cnt > 0x7fffffffLL) return 1;
  buf_put_i32(&c->off[level], (int32_t)cnt);
  return 0;
}
static void colb_free(colb_t* c) {
  buf_free(&c->valid); for (int l = 0; l < 3; l++) buf_free(&c->off[l]); buf_free(&c->values); free(c->bitmap);
}

/* one leaf element of `f` at index i converted to the column's element type
 * (M/TFRecordDeserializer.scala:74-95,102-108) */
static void put_elem(colb_t* c, const feature_t* f, size_t i) {
  switch (c->elem_type) {
    case TFR_T_INT32: { int32_t v = (int32_t)(uint32_t)(uint64_t)f->ints[i]; buf_put(&c->values, &v, 4); break; }   /* .toInt */
    case TFR_T_INT64: buf_put(&c->values, &f->ints[i], 8); break;
    case TFR_T_FLOAT32: buf_put(&c->values, &f->floats[i], 4); break;
    case TFR_T_FLOAT64: case TFR_T_DECIMAL: { double d = (double)f->floats[i]; buf_put(&c->values, &d, 8); break; }  /* .toDouble */
    case TFR_T_STRING: java_utf8_roundtrip(f->bytes[i].p, f->bytes[i].n, &c->values); break;
    case TFR_T_BINARY: buf_put(&c->values, f->bytes[i].p, f->bytes[i].n); break;
    default: break;
  }
}
static int expected_kind(int t) {
  switch (t) { case TFR_T_INT32: case TFR_T_INT64: return K_INT64;
               case TFR_T_FLOAT32: case TFR_T_FLOAT64: case TFR_T_DECIMAL: return K_FLOAT;
               case TFR_T_STRING: case TFR_T_BINARY: return K_BYTES; default: return K_NONE; }
}

/* newFeatureWriter(dataType)(ordinal, feature) for a value of nesting `depth` written at
 * offsets level `level` of column c.  Returns TFR_OK or the error the reference throws. */
static int write_feature(colb_t* c, const feature_t* f, int depth, int level) {
  if (depth == 0) {
    if (c->elem_type == TFR_T_NULL) return 1;   /* caller handles null */
    if (f->kind != expected_kind(c->elem_type)) return TFR_E_KIND_MISMATCH;   /* require(...) :178,189,201,212 */
    if (f->n == 0) return TFR_E_EMPTY_SCALAR;                                  /* .head on empty Seq */
    put_elem(c, f, 0);
    if (is_varlen(c->elem_type) && colb_close(c, level)) return TFR_E_BATCH_TOO_LARGE;
    return TFR_OK;
  }
  if (depth == 1) {                                                             /* :97-117 */
    if (f->kind != expected_kind(c->elem_type)) return TFR_E_KIND_MISMATCH;
    for (size_t i = 0; i < f->n; i++) {
      put_elem(c, f, i);
      if (is_varlen(c->elem_type) && colb_close(c, level + 1)) return TFR_E_BATCH_TOO_LARGE;
    }
    if (colb_close(c, level)) return TFR_E_BATCH_TOO_LARGE;
    return TFR_OK;
  }
  return TFR_E_BAD_NESTING;   /* ArrayType(ArrayType) from a Feature: RuntimeException :119 */
}
/* newFeatureListWriter (:129-143): dataType must be ArrayType(elementType) */
static int write_featurelist(colb_t* c, const featurelist_t* l) {
  if (c->depth == 0) return TFR_E_BAD_NESTING;                                  /* :142 */
  for (size_t i = 0; i < l->n; i++) {
    int rc = write_feature(c, &l->f[i], c->depth - 1, 1);
    if (rc) return rc;
  }
  if (colb_close(c, 0)) return TFR_E_BATCH_TOO_LARGE;
  return TFR_OK;
}

typedef struct tfr_oracle_batch {
  int n_cols; colb_t* cols; tfr_column* view; tfr_batch_info info;
} tfr_oracle_batch;

static void truncate_col(colb_t* c, int64_t rows, size_t* off_n, size_t values_n, size_t valid_n) {
  (void)rows;
  for (int l = 0; l < c->n_levels; l++) c->off[l].n = off_n[l];
  c->values.n = values_n; c->valid.n = valid_n;
}

/* TFRecordReader.read + Example.parseFrom + TFRecordDeserializer.deserialize*, one record at
 * a time in file order (M/TFRecordFileReader.scala:49-81).                                 */
int32_t tfr_oracle_decode(const uint8_t* data, size_t nbytes, const tfr_field* fields, int32_t nf,
                          int32_t record_type, uint32_t flags, int32_t is_final, tfr_oracle_batch** out) {
  crc_init();
  if (record_type < 0 || record_type > 2) return TFR_E_BAD_RECORD_TYPE;
  tfr_oracle_batch* b = (tfr_oracle_batch*)calloc(1, sizeof *b);
  int ncols = record_type == TFR_RT_BYTE_ARRAY ? 1 : nf;
  b->n_cols = ncols; b->cols = (colb_t*)calloc(ncols ? ncols : 1, sizeof(colb_t));
  if (record_type == TFR_RT_BYTE_ARRAY) colb_init(&b->cols[0], TFR_T_BINARY, 0);
  else for (int i = 0; i < nf; i++) colb_init(&b->cols[i], fields[i].elem_type, fields[i].depth);
  b->info.error_row = -1; b->info.error_field = -1;
  size_t pos = 0; int64_t nrec = 0;
  record_t rec; memset(&rec, 0, sizeof rec); rec.feature_lists.is_flist = 1;
  while (1) {
    /* ---- TFRecordReader.read() ---- */
    size_t left = nbytes - pos;
    if (left < 8) {               /* readFully(lenBytes) hits EOF -> EOFException is caught -> null = end */
      if (!is_final) break;       /* non-final block: the fragment is carried over             */
      pos = nbytes; break;        /* final: 0..7 stray bytes are silently dropped by the reader */
    }
    uint64_t len = 0; for (int i = 0; i < 8; i++) len |= (uint64_t)data[pos + i] << (8 * i);
    int err = 0;
    if (left < 12) { if (!is_final) break; err = TFR_E_TRUNCATED; }
    if (!err && (flags & TFR_F_VERIFY_CRC)) {
      uint32_t c = (uint32_t)data[pos + 8] | (uint32_t)data[pos + 9] << 8 | (uint32_t)data[pos + 10] << 16 | (uint32_t)data[pos + 11] << 24;
      if (c != tfr_oracle_masked_crc32c(data + pos, 8)) err = TFR_E_CRC_LENGTH;
    }
    if (!err && len > 0x7fffffffULL) err = TFR_E_RECORD_TOO_LARGE;
    if (!err && left < 16 + len) { if (!is_final) break; err = TFR_E_TRUNCATED; }
    const uint8_t* payload = data + pos + 12;
    if (!err && (flags & TFR_F_VERIFY_CRC)) {
      const uint8_t* q = payload + len;
      uint32_t c = (uint32_t)q[0] | (uint32_t)q[1] << 8 | (uint32_t)q[2] << 16 | (uint32_t)q[3] << 24;
      if (c != tfr_oracle_masked_crc32c(payload, (size_t)len)) err = TFR_E_CRC_DATA;
    }
    int err_field = -1;
    /* remember sizes so that a failing row leaves no partial output behind */
    size_t save_off[64][3]; size_t save_val[64], save_valid[64];
    size_t(*soff)[3] = save_off; size_t *sval = save_val, *svalid = save_valid;
    if (ncols > 64) { soff = (size_t(*)[3])malloc(sizeof(size_t[3]) * ncols); sval = (size_t*)malloc(sizeof(size_t) * ncols); svalid = (size_t*)malloc(sizeof(size_t) * ncols); }
    for (int i = 0; i < ncols; i++) { for (int l = 0; l < 3; l++) soff[i][l] = b->cols[i].off[l].n; sval[i] = b->cols[i].values.n; svalid[i] = b->cols[i].valid.n; }
    if (!err) {
      if (record_type == TFR_RT_BYTE_ARRAY) {         /* deserializeByteArray :17-19 */
        colb_t* c = &b->cols[0];
        buf_put(&c->values, payload, (size_t)len);
        if (colb_close(c, 0)) err = TFR_E_BATCH_TOO_LARGE;
        buf_put_u8(&c->valid, 1);
      } else {
        map_clear(&rec.features); map_clear(&rec.feature_lists);
        if (parse_record(payload, (size_t)len, record_type, &rec)) err = TFR_E_MALFORMED_PROTO;
        /* deserializeExample :21-35 / deserializeSequenceExample :37-61 */
        for (int i = 0; i < nf && !err; i++) {
          colb_t* c = &b->cols[i];
          mapent_t* e = map_get(&rec.features, fields[i].name, (uint32_t)fields[i].name_len);
          int rc = 1;   /* 1 = null */
          if (e) {
            rc = write_feature(c, &e->feat, c->depth, 0);
          } else if (record_type == TFR_RT_SEQUENCE_EXAMPLE &&
                     (e = map_get(&rec.feature_lists, fields[i].name, (uint32_t)fields[i].name_len))) {
            rc = write_featurelist(c, &e->flist);
          } else if (!fields[i].nullable) rc = TFR_E_NULL_IN_NONNULL;           /* :31,56 */
          if (rc < 0) { err = rc; err_field = i; break; }
          if (rc == 1) {   /* null slot: empty extents at every level */
            if (c->n_levels > 0 && colb_close(c, 0)) { err = TFR_E_BATCH_TOO_LARGE; break; }
            if (c->n_levels == 0 && c->width) { uint64_t z = 0; buf_put(&c->values, &z, c->width); }
            buf_put_u8(&c->valid, 0); c->null_count++;
          } else buf_put_u8(&c->valid, 1);
        }
      }
    }
    if (err) {
      for (int i = 0; i < ncols; i++) {
        size_t nv = svalid[i];
        /* recount nulls that were added for this partial row */
        for (size_t k = nv; k < b->cols[i].valid.n; k++) if (!b->cols[i].valid.p[k]) b->cols[i].null_count--;
        truncate_col(&b->cols[i], nrec, soff[i], sval[i], svalid[i]);
      }
      b->info.error_code = err; b->info.error_row = nrec; b->info.error_field = err_field;
      if (ncols > 64) { free(soff); free(sval); free(svalid); }
      break;
    }
    if (ncols > 64) { free(soff); free(sval); free(svalid); }
    pos += 16 + (size_t)len; nrec++;
  }
  map_clear(&rec.features); map_clear(&rec.feature_lists);
  b->info.n_rows = nrec; b->info.n_records = nrec; b->info.consumed_bytes = (int64_t)pos;
  /* finish: bitmaps + views */
  b->view = (tfr_column*)calloc(ncols ? ncols : 1, sizeof(tfr_column));
  int64_t total = 0;
  for (int i = 0; i < ncols; i++) {
    colb_t* c = &b->cols[i]; c->n_rows = nrec;
    size_t nb = (size_t)((nrec + 7) / 8);
    c->bitmap = (uint8_t*)calloc(nb ? nb : 1, 1);
    for (int64_t r = 0; r < nrec; r++) if (c->valid.p[r]) c->bitmap[r >> 3] |= (uint8_t)(1u << (r & 7));
    tfr_column* v = &b->view[i];
    v->elem_type = c->elem_type; v->depth = c->depth; v->n_levels = c->n_levels; v->value_width = c->width;
    v->n_rows = nrec; v->null_count = c->null_count; v->validity = c->bitmap;
    total += (int64_t)nb;
    for (int l = 0; l < c->n_levels; l++) { v->offsets[l] = (int32_t*)c->off[l].p; v->n_offsets[l] = (int64_t)(c->off[l].n / 4); total += (int64_t)c->off[l].n; }
    v->values = c->values.p; v->n_values = c->width ? (int64_t)(c->values.n / c->width) : 0;
    total += (int64_t)c->values.n;
  }
  b->info.out_bytes = total;
  *out = b;
  return TFR_OK;
}
int32_t tfr_oracle_batch_info(tfr_oracle_batch* b, tfr_batch_info* out) { *out = b->info; return TFR_OK; }
int32_t tfr_oracle_batch_columns(tfr_oracle_batch* b, tfr_column* out, int32_t n) {
  if (n < b->n_cols) return TFR_E_INVALID_ARG;
  memcpy(out, b->view, sizeof(tfr_column) * b->n_cols); return TFR_OK;
}
void tfr_oracle_batch_free(tfr_oracle_batch* b) {
  if (!b) return;
  for (int i = 0; i < b->n_cols; i++) colb_free(&b->cols[i]);
  free(b->cols); free(b->view); free(b);
}

/* ======================================================================================
 * encode: TFRecordSerializer.serialize* -> toByteArray -> TFRecordWriter.write
 * (M/TFRecordSerializer.scala:20-60,68-207; M/TFRecordOutputWriter.scala:26-38)
 * ====================================================================================== */
static size_t varint_size(uint64_t v) { size_t n = 1; while (v >= 0x80) { v >>= 7; n++; } return n; }
static void put_varint(buf_t* b, uint64_t v) { while (v >= 0x80) { buf_put_u8(b, (uint8_t)(v | 0x80)); v >>= 7; } buf_put_u8(b, (uint8_t)v); }

/* Serialised bytes of one Feature built from leaf range [lo,hi) of column c
 * (Int64ListFeature/floatListFeature/bytesListFeature :182-207 then Feature.toByteArray):
 * packed repeated fields are omitted when empty, the oneof member is always written.      */
static void emit_feature(buf_t* o, const tfr_column* c, int64_t lo, int64_t hi) {
  buf_t list = {0};
  int kind = expected_kind(c->elem_type);
  if (kind == K_INT64) {
    if (hi > lo) {
      buf_t pk = {0};
      for (int64_t i = lo; i < hi; i++) {
        int64_t v = c->elem_type == TFR_T_INT32 ? (int64_t)((const int32_t*)c->values)[i] : ((const int64_t*)c->values)[i];
        put_varint(&pk, (uint64_t)v);
      }
      buf_put_u8(&list, 0x0A); put_varint(&list, pk.n); buf_put(&list, pk.p, pk.n); buf_free(&pk);
    }
    buf_put_u8(o, 0x1A);
  } else if (kind == K_FLOAT) {
    if (hi > lo) {
      buf_put_u8(&list, 0x0A); put_varint(&list, (uint64_t)(hi - lo) * 4);
      for (int64_t i = lo; i < hi; i++) {
        float f = c->elem_type == TFR_T_FLOAT32 ? ((const float*)c->values)[i] : (float)((const double*)c->values)[i]; /* toFloat :86,113 */
        buf_put(&list, &f, 4);
      }
    }
    buf_put_u8(o, 0x12);
  } else {
    const int32_t* so = c->offsets[c->n_levels - 1];
    for (int64_t i = lo; i < hi; i++) {
      uint32_t l = (uint32_t)(so[i + 1] - so[i]);
      buf_put_u8(&list, 0x0A); put_varint(&list, l); buf_put(&list, (const uint8_t*)c->values + so[i], l);
    }
    buf_put_u8(o, 0x0A);
  }
  put_varint(o, list.n); buf_put(o, list.p, list.n); buf_free(&list);
}
static void emit_entry(buf_t* o, const tfr_field* f, const buf_t* val) {
  size_t elen = 1 + varint_size((uint64_t)f->name_len) + (size_t)f->name_len + 1 + varint_size(val->n) + val->n;
  buf_put_u8(o, 0x0A); put_varint(o, elen);
  buf_put_u8(o, 0x0A); put_varint(o, (uint64_t)f->name_len); buf_put(o, f->name, (size_t)f->name_len);
  buf_put_u8(o, 0x12); put_varint(o, val->n); buf_put(o, val->p, val->n);
}
static int bit_get(const uint8_t* bm, int64_t i) { return bm ? (bm[i >> 3] >> (i & 7)) & 1 : 1; }

int32_t tfr_oracle_encode(const tfr_field* fields, int32_t nf, int32_t record_type, const tfr_column* cols,
                          int64_t n_rows, uint8_t** out, size_t* out_bytes, int64_t* error_row) {
  crc_init();
  buf_t file = {0};
  *error_row = -1;
  for (int64_t r = 0; r < n_rows; r++) {
    buf_t rec = {0};
    if (record_type == TFR_RT_BYTE_ARRAY) {            /* serializeByteArray :16-18 */
      const tfr_column* c = &cols[0];
      buf_put(&rec, (const uint8_t*)c->values + c->offsets[0][r], (size_t)(c->offsets[0][r + 1] - c->offsets[0][r]));
    } else {
      buf_t ctx = {0}, fl = {0};
      for (int i = 0; i < nf; i++) {
        const tfr_column* c = &cols[i];
        if (!bit_get(c->validity, r) || c->elem_type == TFR_T_NULL) {
          if (!fields[i].nullable) { *error_row = r; buf_free(&ctx); buf_free(&fl); buf_free(&rec); buf_free(&file); return TFR_E_NULL_IN_NONNULL; }
          continue;                                     /* :25-31 null + nullable -> omitted */
        }
        buf_t val = {0};
        if (c->depth == 2) {                            /* FeatureList :138-145 */
          const int32_t* o0 = c->offsets[0]; const int32_t* o1 = c->offsets[1];
          for (int32_t s = o0[r]; s < o0[r + 1]; s++) {
            buf_t ft = {0}; emit_feature(&ft, c, o1[s], o1[s + 1]);
            buf_put_u8(&val, 0x0A); put_varint(&val, ft.n); buf_put(&val, ft.p, ft.n); buf_free(&ft);
          }
          emit_entry(&fl, &fields[i], &val);
        } else {
          int64_t lo = r, hi = r + 1;
          if (c->depth == 1) { lo = c->offsets[0][r]; hi = c->offsets[0][r + 1]; }
          emit_feature(&val, c, lo, hi);
          emit_entry(&ctx, &fields[i], &val);
        }
        buf_free(&val);
      }
      /* setFeatures / setContext + setFeatureLists are always called (:33,57-58) */
      buf_put_u8(&rec, 0x0A); put_varint(&rec, ctx.n); buf_put(&rec, ctx.p, ctx.n);
      if (record_type == TFR_RT_SEQUENCE_EXAMPLE) { buf_put_u8(&rec, 0x12); put_varint(&rec, fl.n); buf_put(&rec, fl.p, fl.n); }
      buf_free(&ctx); buf_free(&fl);
    }
    /* TFRecordWriter.write */
    uint8_t hdr[12]; uint64_t len = rec.n;
    for (int i = 0; i < 8; i++) hdr[i] = (uint8_t)(len >> (8 * i));
    uint32_t c1 = tfr_oracle_masked_crc32c(hdr, 8); for (int i = 0; i < 4; i++) hdr[8 + i] = (uint8_t)(c1 >> (8 * i));
    buf_put(&file, hdr, 12); buf_put(&file, rec.p, rec.n);
    uint32_t c2 = tfr_oracle_masked_crc32c(rec.p ? rec.p : (const uint8_t*)"", rec.n);
    uint8_t ft[4]; for (int i = 0; i < 4; i++) ft[i] = (uint8_t)(c2 >> (8 * i));
    buf_put(&file, ft, 4);
    buf_free(&rec);
  }
  *out = file.p; *out_bytes = file.n;
  if (!file.p) *out = (uint8_t*)malloc(1);
  return TFR_OK;
}
void tfr_oracle_free(void* p) { free(p); }

/* ======================================================================================
 * schema inference (M/TensorFlowInferSchema.scala:35-228): per record name -> lattice code,
 * merged with findTightestCommonType (= max with 0/null as identity, :213-228)
 * ====================================================================================== */
typedef struct { char* name; uint32_t len; int code; } infent_t;
typedef struct tfr_oracle_infer_t { infent_t* e; size_t n, cap; } tfr_oracle_infer_t;
static int infer_feature_code(const feature_t* f) {     /* inferField :132-145 + parse*List :147-188 */
  if (f->kind == K_NONE) return -1;                     /* KIND_NOT_SET -> exception :143 */
  if (f->n == 0) return TFR_INF_NULL;
  int base = f->kind == K_INT64 ? TFR_INF_LONG : f->kind == K_FLOAT ? TFR_INF_FLOAT : TFR_INF_STRING;
  return f->n > 1 ? base + 3 : base;
}
static int infer_conflict = 0;
static void infer_merge(tfr_oracle_infer_t* s, const char* name, uint32_t len, int code) {
  for (size_t i = 0; i < s->n; i++)
    if (s->e[i].len == len && memcmp(s->e[i].name, name, len) == 0) {
      /* findTightestCommonType: equal -> same; null is the identity; ArrayType(ArrayType(null)) has no precedence -> throws */
      int old = s->e[i].code;
      if (old != code && old != TFR_INF_NULL && code != TFR_INF_NULL && (old == TFR_INF_ARR2_NULL || code == TFR_INF_ARR2_NULL)) infer_conflict = 1;
      if (code > old) s->e[i].code = code;
      return;
    }
  if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 16; s->e = (infent_t*)realloc(s->e, s->cap * sizeof(infent_t)); }
  s->e[s->n].name = (char*)malloc(len ? len : 1); memcpy(s->e[s->n].name, name, len);
  s->e[s->n].len = len; s->e[s->n].code = code; s->n++;
}
int32_t tfr_oracle_infer(const uint8_t* data, size_t nbytes, int32_t record_type, tfr_oracle_infer_t** out) {
  tfr_oracle_infer_t* s = (tfr_oracle_infer_t*)calloc(1, sizeof *s);
  infer_conflict = 0;
  record_t rec; memset(&rec, 0, sizeof rec); rec.feature_lists.is_flist = 1;
  size_t pos = 0; int rc = TFR_OK;
  while (nbytes - pos >= 8) {
    uint64_t len = 0; for (int i = 0; i < 8; i++) len |= (uint64_t)data[pos + i] << (8 * i);
    if (nbytes - pos < 16 + len) { rc = TFR_E_TRUNCATED; break; }
    map_clear(&rec.features); map_clear(&rec.feature_lists);
    if (parse_record(data + pos + 12, (size_t)len, record_type, &rec)) { rc = TFR_E_MALFORMED_PROTO; break; }
    for (size_t i = 0; i < rec.features.n && !rc; i++) {
      int c = infer_feature_code(&rec.features.e[i].feat);
      if (c < 0) { rc = TFR_E_KIND_MISMATCH; break; }
      infer_merge(s, rec.features.e[i].key, rec.features.e[i].key_len, c);
    }
    for (size_t i = 0; i < rec.feature_lists.n && !rc; i++) {   /* inferFeatureListTypes :98-118 */
      featurelist_t* l = &rec.feature_lists.e[i].flist; int c = TFR_INF_NULL;
      if (l->n == 0) { rc = TFR_E_EMPTY_SCALAR; break; }          /* empty.reduceLeft -> UnsupportedOperationException */
      for (size_t k = 0; k < l->n; k++) {
        int ck = infer_feature_code(&l->f[k]);
        if (ck < 0) { rc = TFR_E_KIND_MISMATCH; break; }
        if (ck > c) c = ck;                                          /* reduceLeft(findTightestCommonType) */
      }
      if (rc) break;
      /* T or ArrayType(T) -> ArrayType(ArrayType(T)) (:102-107); all steps empty -> ArrayType(ArrayType(null)) = 10 */
      if (c != TFR_INF_NULL) { int base = (c - 1) % 3; c = TFR_INF_ARR2_LONG + base; } else c = TFR_INF_ARR2_NULL;
      infer_merge(s, rec.feature_lists.e[i].key, rec.feature_lists.e[i].key_len, c);
    }
    if (rc) break;
    pos += 16 + (size_t)len;
  }
  map_clear(&rec.features); map_clear(&rec.feature_lists);
  *out = s;
  if (!rc && infer_conflict) rc = TFR_E_UNSUPPORTED_TYPE;
  return rc;
}
int32_t tfr_oracle_infer_count(tfr_oracle_infer_t* s) { return (int32_t)s->n; }
int32_t tfr_oracle_infer_get(tfr_oracle_infer_t* s, int32_t i, const char** name, int32_t* len, int32_t* code) {
  *name = s->e[i].name; *len = (int32_t)s->e[i].len; *code = s->e[i].code; return TFR_OK;
}
void tfr_oracle_infer_free(tfr_oracle_infer_t* s) { if (!s) return; for (size_t i = 0; i < s->n; i++) free(s->e[i].name); free(s->e); free(s); }
