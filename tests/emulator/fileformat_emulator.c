/*
 * fileformat_emulator.c -- plays, in plain C against include/tfrgpu.h ONLY, the call sequence Spark drives through the
 * reference's FileFormat for the hot path (no Python, no torch, no oracle):
 *
 *   write side  M/DefaultSource.scala:105-110 (OutputWriterFactory.newInstance), M/TFRecordOutputWriter.scala:12-24 (ctor),
 *               :26-38 (write(row), once per row), :40-43 (close(), exactly once)
 *   read side   M/DefaultSource.scala:118-136 (buildReader closure, once per PartitionedFile),
 *               M/TFRecordFileReader.scala:16-44 (setup), :46-82 (hasNext/next pulled row by row), :36-40,52-57 (idempotent
 *               close: at EOF by hasNext, again by the task-completion listener)
 *
 * What a JNI shim does between those calls and libtfrgpu.so is exactly what this file does: stage file blocks in the
 * decoder's pinned slots, tfr_decode_submit the next block while rows of the current one are consumed, carry the
 * unconsumed tail, deliver rows before an error and then fail like the reference's iterator, buffer rows column-wise and
 * tfr_encode them at every flush and at close.
 *
 *   fileformat_emulator abi                      -> ABI version + every status string (runs without a GPU)
 *   fileformat_emulator roundtrip DIR N BLOCK    -> write N rows, read them back in BLOCK-byte blocks, compare; then corrupt
 *                                                   one record and check rows-before-error + the error class
 * Exit code 0 = every check passed.  Rows are a deterministic function of the row index (row_value_* below), so the
 * Python side of the test can regenerate them and compare the file with the CPU oracle's reading of it.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tfrgpu.h"

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "emulator: %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } } while (0)
#define OK(call) do { int32_t rc_ = (call); CHECK(rc_ == 0, "%s -> %d (%s: %s)", #call, rc_, tfr_status_string(rc_), tfr_last_error()); } while (0)

/* ---- the data schema of this emulation: StructType(id: Long, w: Float, name: String (nullable), emb: Array[Float]) ---- */
enum { N_FIELDS = 4 };
static const tfr_field FIELDS[N_FIELDS] = {
  {"id", 2, TFR_T_INT64, 0, 0}, {"w", 1, TFR_T_FLOAT32, 0, 1}, {"name", 4, TFR_T_STRING, 0, 1}, {"emb", 3, TFR_T_FLOAT32, 1, 1},
};
static int64_t row_value_id(int64_t i) { return i * i - 7 * i - 3; }
static float row_value_w(int64_t i) { return (float)i * 0.5f - 100.0f; }
static int row_name_is_null(int64_t i) { return i % 11 == 5; }
static int row_value_name(int64_t i, char* out) { return sprintf(out, "row-%lld-%s", (long long)i, (i % 3) ? "x" : "yy"); }
static int row_emb_len(int64_t i) { return (int)(i % 6); }
static float row_value_emb(int64_t i, int k) { return (float)(i + k) * 0.25f; }

static uint64_t fnv(uint64_t h, const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } return h; }
static uint64_t row_hash(uint64_t h, int64_t id, float w, int name_null, const char* name, int name_len, int emb_len, const float* emb) {
  h = fnv(h, &id, 8); h = fnv(h, &w, 4); h = fnv(h, &name_null, 4);
  if (!name_null) h = fnv(h, name, (size_t)name_len);
  h = fnv(h, &emb_len, 4); h = fnv(h, emb, 4 * (size_t)emb_len);
  return h;
}

/* ================================ OutputWriter ================================ */
typedef struct {
  tfr_schema* schema; tfr_encoder* enc; FILE* out; int closed;
  int64_t n, cap;                      /* buffered rows */
  int64_t* id; float* w; uint8_t* v_w; uint8_t* v_name; uint8_t* v_all;
  int32_t* name_off; char* name_bytes; int64_t name_cap;
  int32_t* emb_off; float* emb_vals; int64_t emb_cap;
  int64_t flush_rows, rows_written, flushes;
} Writer;

static void writer_new_instance(Writer* W, const char* path, int64_t flush_rows) {   /* OutputWriterFactory.newInstance -> ctor */
  memset(W, 0, sizeof *W);
  OK(tfr_schema_create(FIELDS, N_FIELDS, TFR_RT_EXAMPLE, &W->schema));
  OK(tfr_encoder_create(W->schema, 0, 0, &W->enc));
  W->out = fopen(path, "wb");
  CHECK(W->out, "cannot create %s", path);
  W->flush_rows = flush_rows; W->cap = flush_rows;
  W->id = malloc(8 * W->cap); W->w = malloc(4 * W->cap);
  W->v_w = calloc((W->cap + 7) / 8, 1); W->v_name = calloc((W->cap + 7) / 8, 1); W->v_all = malloc((W->cap + 7) / 8);
  memset(W->v_all, 0xFF, (W->cap + 7) / 8);
  W->name_off = malloc(4 * (W->cap + 1)); W->name_cap = 64 * W->cap; W->name_bytes = malloc(W->name_cap);
  W->emb_off = malloc(4 * (W->cap + 1)); W->emb_cap = 8 * W->cap; W->emb_vals = malloc(4 * W->emb_cap);
  W->name_off[0] = 0; W->emb_off[0] = 0;
}
static void writer_flush(Writer* W) {
  if (!W->n) return;
  tfr_column c[N_FIELDS];
  memset(c, 0, sizeof c);
  for (int f = 0; f < N_FIELDS; ++f) { c[f].elem_type = FIELDS[f].elem_type; c[f].depth = FIELDS[f].depth; c[f].n_rows = W->n; }
  c[0].value_width = 8; c[0].validity = W->v_all; c[0].values = W->id; c[0].n_values = W->n;
  c[1].value_width = 4; c[1].validity = W->v_w; c[1].values = W->w; c[1].n_values = W->n;
  c[2].value_width = 1; c[2].n_levels = 1; c[2].validity = W->v_name; c[2].offsets[0] = W->name_off; c[2].n_offsets[0] = W->n + 1;
  c[2].values = W->name_bytes; c[2].n_values = W->name_off[W->n];
  c[3].value_width = 4; c[3].n_levels = 1; c[3].validity = W->v_all; c[3].offsets[0] = W->emb_off; c[3].n_offsets[0] = W->n + 1;
  c[3].values = W->emb_vals; c[3].n_values = W->emb_off[W->n];
  void* dev = NULL; void* host = NULL; size_t nb = 0; int64_t err_row = -1;
  OK(tfr_encode(W->enc, c, N_FIELDS, 0, &dev, &nb, &err_row));
  OK(tfr_encoder_result_host(W->enc, &host, &nb));
  CHECK(fwrite(host, 1, nb, W->out) == nb, "short write");          /* outputStream.write: the bytes the reference writer appends */
  W->rows_written += W->n; W->flushes++;
  memset(W->v_w, 0, (W->cap + 7) / 8); memset(W->v_name, 0, (W->cap + 7) / 8);
  W->n = 0;
}
static void writer_write(Writer* W, int64_t i) {                     /* OutputWriter.write(row: InternalRow) */
  CHECK(!W->closed, "write after close");
  const int64_t r = W->n;
  W->id[r] = row_value_id(i);
  W->w[r] = row_value_w(i); W->v_w[r >> 3] |= (uint8_t)(1u << (r & 7));
  if (!row_name_is_null(i)) {
    char tmp[64]; int l = row_value_name(i, tmp);
    memcpy(W->name_bytes + W->name_off[r], tmp, (size_t)l);
    W->name_off[r + 1] = W->name_off[r] + l; W->v_name[r >> 3] |= (uint8_t)(1u << (r & 7));
  } else W->name_off[r + 1] = W->name_off[r];
  int el = row_emb_len(i);
  for (int k = 0; k < el; ++k) W->emb_vals[W->emb_off[r] + k] = row_value_emb(i, k);
  W->emb_off[r + 1] = W->emb_off[r] + el;
  if (++W->n == W->flush_rows) writer_flush(W);
}
static void writer_close(Writer* W) {                                /* OutputWriter.close(): Spark calls it exactly once */
  CHECK(!W->closed, "OutputWriter.close called twice");
  writer_flush(W);
  fclose(W->out);
  tfr_encoder_destroy(W->enc); tfr_schema_destroy(W->schema);
  free(W->id); free(W->w); free(W->v_w); free(W->v_name); free(W->v_all); free(W->name_off); free(W->name_bytes); free(W->emb_off); free(W->emb_vals);
  W->closed = 1;
}

/* ================================ buildReader closure ================================ */
typedef struct {
  tfr_schema* schema; tfr_decoder* dec; FILE* in; int closed, n_close_calls;
  size_t block; int64_t remaining;
  /* pipeline: block k is staged in slot k % slots; at most one block is submitted ahead of the one being iterated */
  int slots, next_slot;
  uint8_t* carry; size_t carry_len, carry_cap;
  tfr_batch* cur; tfr_column cols[N_FIELDS]; int64_t cur_row, cur_rows; int32_t cur_err; int64_t cur_err_row;
  tfr_batch* ahead; int ahead_final, ahead_slot; size_t ahead_nbytes;
  int eof_submitted, finished;
  int64_t rows_delivered, blocks;
} Reader;

static void reader_close(Reader* R) {                                /* M/TFRecordFileReader.scala:36-40: safe to call again */
  R->n_close_calls++;
  if (R->closed) return;
  if (R->cur) tfr_batch_release(R->cur);
  if (R->ahead) tfr_batch_release(R->ahead);
  R->cur = R->ahead = NULL;
  tfr_decoder_destroy(R->dec); tfr_schema_destroy(R->schema);
  if (R->in) fclose(R->in);
  free(R->carry);
  R->closed = 1;
}
/* stage [carry | next file bytes] in a pinned slot and submit it; the tail a previous block left over is known by now */
static void reader_submit_next(Reader* R) {
  if (R->eof_submitted) return;
  const int slot = R->next_slot; R->next_slot = (R->next_slot + 1) % R->slots;
  size_t want = R->block > R->carry_len ? R->block - R->carry_len : R->block / 2 + 1;   /* a carried record larger than the block still makes progress */
  if ((int64_t)want > R->remaining) want = (size_t)R->remaining;
  void* st = NULL; size_t cap = 0;
  OK(tfr_decoder_staging_slot(R->dec, slot, R->carry_len + want + 16, &st, &cap));
  memcpy(st, R->carry, R->carry_len);
  size_t got = want ? fread((uint8_t*)st + R->carry_len, 1, want, R->in) : 0;
  R->remaining -= (int64_t)got;
  const int final = R->remaining == 0 || got < want;
  const size_t nbytes = R->carry_len + got;
  OK(tfr_decode_submit(R->dec, st, nbytes, 0, final, &R->ahead));
  R->ahead_final = final; R->ahead_slot = slot; R->ahead_nbytes = nbytes;
  R->blocks++;
  if (final) R->eof_submitted = 1;
}
static void reader_open(Reader* R, const char* path, size_t block) { /* buildReader(...)(file) -> TFRecordFileReader.readFile */
  memset(R, 0, sizeof *R);
  OK(tfr_schema_create(FIELDS, N_FIELDS, TFR_RT_EXAMPLE, &R->schema));   /* requiredSchema reaches the parser (M/DefaultSource.scala:134) */
  OK(tfr_decoder_create(R->schema, 0, TFR_F_DEFAULT, &R->dec));
  R->in = fopen(path, "rb");
  CHECK(R->in, "cannot open %s", path);
  fseek(R->in, 0, SEEK_END); R->remaining = ftell(R->in); fseek(R->in, 0, SEEK_SET);
  R->block = block; R->slots = tfr_decoder_num_staging_slots();
  CHECK(R->slots >= 2, "need two staging slots to read ahead");
  R->carry_cap = 1 << 16; R->carry = malloc(R->carry_cap);
  reader_submit_next(R);
}
/* make the submitted block the current one.  Where it ends is known as soon as its frame index has run (tfr_batch_consumed):
 * its tail is carried over and the block after it is submitted BEFORE this one's rows are waited for, so the next block's
 * copy and frame index overlap this block's decode and the consumption of its rows */
static void reader_advance(Reader* R) {
  if (R->cur) { tfr_batch_release(R->cur); R->cur = NULL; }
  if (!R->ahead) { R->finished = 1; return; }
  tfr_batch* b = R->ahead; R->ahead = NULL;
  if (!R->ahead_final) {
    size_t used = 0;
    OK(tfr_batch_consumed(b, &used));
    void* st = NULL; size_t cap = 0;
    OK(tfr_decoder_staging_slot(R->dec, R->ahead_slot, 0, &st, &cap));
    R->carry_len = R->ahead_nbytes - used;                            /* the partial record sits in the staging slot behind `used` */
    if (R->carry_len > R->carry_cap) { R->carry_cap = R->carry_len * 2; R->carry = realloc(R->carry, R->carry_cap); }
    memcpy(R->carry, (uint8_t*)st + used, R->carry_len);
    reader_submit_next(R);
  }
  tfr_batch_info info;
  OK(tfr_batch_status(b, &info));                                    /* waits for the rows; with an error, consumed_bytes is the final word and the stream ends here */
  OK(tfr_batch_to_host(b, R->cols, N_FIELDS));
  R->cur = b; R->cur_row = 0; R->cur_rows = info.n_rows; R->cur_err = info.error_code; R->cur_err_row = info.error_row;
}
/* Iterator.hasNext / next fused: 1 = a row was delivered into *h, 0 = end of file, < 0 = the status the reference would throw for */
static int32_t reader_next(Reader* R, uint64_t* h) {
  for (;;) {
    if (R->finished) return 0;
    if (!R->cur) { reader_advance(R); continue; }
    if (R->cur_row < R->cur_rows) {
      const int64_t r = R->cur_row++;
      const tfr_column* c = R->cols;
      const int64_t id = ((const int64_t*)c[0].values)[r];
      const int w_valid = (c[1].validity[r >> 3] >> (r & 7)) & 1;
      const float w = w_valid ? ((const float*)c[1].values)[r] : 0.0f;
      const int name_null = !((c[2].validity[r >> 3] >> (r & 7)) & 1);
      const int32_t n0 = c[2].offsets[0][r], n1 = c[2].offsets[0][r + 1];
      const int32_t e0 = c[3].offsets[0][r], e1 = c[3].offsets[0][r + 1];
      *h = row_hash(*h, id, w, name_null, (const char*)c[2].values + n0, n1 - n0, e1 - e0, (const float*)c[3].values + e0);
      R->rows_delivered++;
      return 1;
    }
    if (R->cur_err) { const int32_t e = R->cur_err; reader_close(R); R->finished = 1; return e; }   /* next() throws after the rows before the bad record */
    if (!R->ahead) { reader_close(R); R->finished = 1; return 0; }    /* hasNext: EOF closes the reader (:52-57) */
    reader_advance(R);
  }
}

static uint64_t expected_hash(int64_t n) {
  uint64_t h = 1469598103934665603ull;
  for (int64_t i = 0; i < n; ++i) {
    char nm[64]; int nl = 0; float emb[8];
    const int nn = row_name_is_null(i);
    if (!nn) nl = row_value_name(i, nm);
    const int el = row_emb_len(i);
    for (int k = 0; k < el; ++k) emb[k] = row_value_emb(i, k);
    h = row_hash(h, row_value_id(i), row_value_w(i), nn, nm, nl, el, emb);
  }
  return h;
}

static int cmd_abi(void) {
  printf("abi %d\n", tfr_abi_version());
  for (int s = 0; s >= -18; --s) printf("%d %s\n", s, tfr_status_string(s));
  tfr_schema* sch = NULL;
  OK(tfr_schema_create(FIELDS, N_FIELDS, TFR_RT_EXAMPLE, &sch));      /* host-only: no device needed */
  CHECK(tfr_schema_num_fields(sch) == N_FIELDS, "schema field count");
  tfr_field bad = {"t", 1, 99, 0, 1};
  tfr_schema* s2 = NULL;
  CHECK(tfr_schema_create(&bad, 1, TFR_RT_EXAMPLE, &s2) == TFR_E_UNSUPPORTED_TYPE, "unsupported type must be rejected at construction");
  CHECK(tfr_schema_create(FIELDS, N_FIELDS, 7, &s2) == TFR_E_BAD_RECORD_TYPE, "bad recordType");
  tfr_schema_destroy(sch);
  printf("staging slots %d\n", tfr_decoder_num_staging_slots());
  return 0;
}

static int cmd_roundtrip(const char* dir, int64_t n, size_t block) {
  char path[1024], bad_path[1024];
  snprintf(path, sizeof path, "%s/part-00000.tfrecord", dir);
  snprintf(bad_path, sizeof bad_path, "%s/part-00001.tfrecord", dir);
  /* ---- df.write.format("tfrecord").save(dir): one task, one OutputWriter ---- */
  Writer W;
  writer_new_instance(&W, path, 20000);
  for (int64_t i = 0; i < n; ++i) writer_write(&W, i);
  writer_close(&W);
  CHECK(W.rows_written == n, "rows written %lld != %lld", (long long)W.rows_written, (long long)n);
  /* ---- spark.read.format("tfrecord").schema(s).load(dir): the closure is called once for the file, rows are pulled one by one ---- */
  const uint64_t want = expected_hash(n);
  Reader R;
  reader_open(&R, path, block);
  uint64_t h = 1469598103934665603ull;
  int32_t rc;
  while ((rc = reader_next(&R, &h)) == 1) {}
  CHECK(rc == 0, "reader failed with %d (%s)", rc, tfr_status_string(rc));
  CHECK(R.rows_delivered == n, "rows read %lld != %lld", (long long)R.rows_delivered, (long long)n);
  CHECK(h == want, "row contents differ after the round trip");
  CHECK(R.closed && R.n_close_calls == 1, "reader must be closed at EOF");
  reader_close(&R);                                                  /* the task-completion listener closes again: must be harmless */
  printf("roundtrip ok: rows=%lld blocks=%lld flushes=%lld hash=%016llx\n", (long long)n, (long long)R.blocks, (long long)W.flushes, (unsigned long long)h);
  /* ---- a corrupt record: rows before it are delivered, then the IOException-class status ---- */
  if (n >= 1000) {
    FILE* f = fopen(path, "rb"); FILE* g = fopen(bad_path, "wb");
    CHECK(f && g, "copy");
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* buf = malloc((size_t)sz);
    CHECK(fread(buf, 1, (size_t)sz, f) == (size_t)sz, "read");
    /* find record `victim` by walking the length fields, flip a payload bit */
    const int64_t victim = n / 2 + 17;
    size_t pos = 0;
    for (int64_t i = 0; i < victim; ++i) { uint64_t len; memcpy(&len, buf + pos, 8); pos += 16 + (size_t)len; }
    buf[pos + 12 + 3] ^= 0x04;
    CHECK(fwrite(buf, 1, (size_t)sz, g) == (size_t)sz, "write");
    fclose(f); fclose(g); free(buf);
    Reader B;
    reader_open(&B, bad_path, block);
    uint64_t hb = 1469598103934665603ull;
    while ((rc = reader_next(&B, &hb)) == 1) {}
    CHECK(rc == TFR_E_CRC_DATA, "corrupt payload must surface as 'Data crc32 checking failed', got %d", rc);
    CHECK(B.rows_delivered == victim, "rows before the bad record: %lld != %lld", (long long)B.rows_delivered, (long long)victim);
    CHECK(hb == expected_hash(victim), "rows before the bad record differ");
    CHECK(B.closed, "reader closed after the failure");
    reader_close(&B);
    printf("error path ok: %lld rows, then %s\n", (long long)victim, tfr_status_string(rc));
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "abi")) return cmd_abi();
  if (argc >= 5 && !strcmp(argv[1], "roundtrip")) return cmd_roundtrip(argv[2], atoll(argv[3]), (size_t)atoll(argv[4]));
  fprintf(stderr, "usage: %s abi | roundtrip DIR N_ROWS BLOCK_BYTES\n", argv[0]);
  return 2;
}
