// tfrgpu_jni.cpp -- JNI shim a spark-tfrecord maintainer adds next to the Scala sources.
// SOURCE ONLY: this image has no JDK (no jni.h), so this file is not compiled by build(); it is kept
// compile-clean against the JNI specification and exercises exactly the C ABI of include/tfrgpu.h.
//
// Java side (package com.linkedin.spark.datasources.tfrecord):
//   final class TfrGpu {
//     static native long schemaCreate(String[] names, int[] elemTypes, int[] depths, boolean[] nullable, int recordType);
//     static native void schemaDestroy(long schema);
//     static native long decoderCreate(long schema, int device, int flags);
//     static native void decoderDestroy(long decoder);    // TaskCompletionListener + the iterator's idempotent close (M/TFRecordFileReader.scala:36-40,52-57)
//     static native java.nio.ByteBuffer decoderStaging(long decoder, int slot, long minBytes);   // direct, pinned; slots 0..stagingSlots()-1
//     static native int stagingSlots();
//     static native long decode(long decoder, java.nio.ByteBuffer staged, long nbytes, boolean isFinal, long[] consumedOut);
//     static native long decodeSubmit(long decoder, java.nio.ByteBuffer staged, long nbytes, boolean isFinal);   // pipelined: no wait
//     static native void batchToHostAsync(long batch);                                                            // D2H behind the kernels
//     static native long[] batchStatus(long batch);      // {nRows, nRecords, consumed, errorCode, errorRow, errorField}; waits for a submitted batch
//     static native java.nio.ByteBuffer[] batchColumnHost(long batch, int column, long[] meta);  // validity, offsets*, values
//     static native void batchExportArrowDevice(long batch, int column, long arrowDeviceArrayAddr, long arrowSchemaAddr);  // ColumnarBatch on the GPU (spark-rapids)
//     static native void batchThrowIfError(long batch);   // the exception the reference would throw for the first failing record
//     static native void batchRelease(long batch);
//     static native long encoderCreate(long schema, int device);
//     static native void encoderDestroy(long encoder);    // OutputWriter.close (M/TFRecordOutputWriter.scala:40-43)
//     static native java.nio.ByteBuffer encode(long encoder, long[] columnStructAddrs, int n);   // framed bytes, pinned
//     static native long inferCreate(int recordType, int device);                                 // DefaultSource.inferSchema (M/DefaultSource.scala:31-39)
//     static native long inferUpdate(long infer, java.nio.ByteBuffer block, long nbytes, boolean isFinal);   // -> consumed bytes
//     static native Object[] inferResult(long infer);     // {String[] names (bytewise sorted), int[] lattice codes}
//     static native void inferDestroy(long infer);
//   }
#ifdef TFR_BUILD_JNI
#include <jni.h>
#include <string>
#include <vector>
#include "../../include/tfrgpu.h"

static void throw_for(JNIEnv* env, int32_t code, int64_t row) {
  const char* cls = "java/lang/RuntimeException";
  switch (code) {
    case TFR_E_CRC_LENGTH: case TFR_E_CRC_DATA: case TFR_E_TRUNCATED: case TFR_E_RECORD_TOO_LARGE: cls = "java/io/IOException"; break;
    case TFR_E_MALFORMED_PROTO: cls = "com/google/protobuf/InvalidProtocolBufferException"; break;
    case TFR_E_KIND_MISMATCH: case TFR_E_BAD_RECORD_TYPE: cls = "java/lang/IllegalArgumentException"; break;
    case TFR_E_EMPTY_SCALAR: cls = "java/util/NoSuchElementException"; break;
    case TFR_E_NULL_IN_NONNULL: cls = "java/lang/NullPointerException"; break;
    case TFR_E_UNSUPPORTED_TYPE: case TFR_E_BAD_NESTING: cls = "java/lang/RuntimeException"; break;
    default: break;
  }
  std::string msg = std::string(tfr_status_string(code)) + (row >= 0 ? " (record " + std::to_string(row) + ")" : "") + ": " + tfr_last_error();
  env->ThrowNew(env->FindClass(cls), msg.c_str());
}

extern "C" JNIEXPORT jlong JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_schemaCreate(
    JNIEnv* env, jclass, jobjectArray names, jintArray elemTypes, jintArray depths, jbooleanArray nullable, jint recordType) {
  jsize n = env->GetArrayLength(names);
  std::vector<std::string> keep(n);
  std::vector<tfr_field> f(n);
  jint* et = env->GetIntArrayElements(elemTypes, nullptr);
  jint* dp = env->GetIntArrayElements(depths, nullptr);
  jboolean* nl = env->GetBooleanArrayElements(nullable, nullptr);
  for (jsize i = 0; i < n; ++i) {
    jstring s = (jstring)env->GetObjectArrayElement(names, i);
    const char* u = env->GetStringUTFChars(s, nullptr);      // note: modified UTF-8; use String.getBytes(UTF_8) for non-BMP names
    keep[i] = u; env->ReleaseStringUTFChars(s, u);
    f[i] = tfr_field{keep[i].data(), (int32_t)keep[i].size(), et[i], dp[i], nl[i] ? 1 : 0};
  }
  tfr_schema* out = nullptr;
  int32_t rc = tfr_schema_create(f.data(), n, recordType, &out);
  env->ReleaseIntArrayElements(elemTypes, et, JNI_ABORT); env->ReleaseIntArrayElements(depths, dp, JNI_ABORT);
  env->ReleaseBooleanArrayElements(nullable, nl, JNI_ABORT);
  if (rc) { throw_for(env, rc, -1); return 0; }
  return (jlong)out;
}
extern "C" JNIEXPORT jlong JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_decoderCreate(JNIEnv* env, jclass, jlong schema, jint device, jint flags) {
  tfr_decoder* d = nullptr;
  int32_t rc = tfr_decoder_create((const tfr_schema*)schema, device, (uint32_t)flags, &d);
  if (rc) { throw_for(env, rc, -1); return 0; }
  return (jlong)d;
}
extern "C" JNIEXPORT void JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_schemaDestroy(JNIEnv*, jclass, jlong schema) { tfr_schema_destroy((tfr_schema*)schema); }
// Safe while batches are still alive (they hold a reference on the decoder) and safe to call from the task-completion
// listener after the iterator already closed: the Scala side nulls its handle, a 0 handle is a no-op here.
extern "C" JNIEXPORT void JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_decoderDestroy(JNIEnv*, jclass, jlong dec) { if (dec) tfr_decoder_destroy((tfr_decoder*)dec); }
extern "C" JNIEXPORT jint JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_stagingSlots(JNIEnv*, jclass) { return tfr_decoder_num_staging_slots(); }
extern "C" JNIEXPORT jobject JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_decoderStaging(JNIEnv* env, jclass, jlong dec, jint slot, jlong minBytes) {
  void* p = nullptr; size_t cap = 0;
  int32_t rc = tfr_decoder_staging_slot((tfr_decoder*)dec, slot, (size_t)minBytes, &p, &cap);
  if (rc) { throw_for(env, rc, -1); return nullptr; }
  return env->NewDirectByteBuffer(p, (jlong)cap);           // the InputStream is read straight into pinned memory
}
extern "C" JNIEXPORT jlong JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_decode(
    JNIEnv* env, jclass, jlong dec, jobject staged, jlong nbytes, jboolean isFinal, jlongArray consumedOut) {
  void* p = env->GetDirectBufferAddress(staged);
  tfr_batch* b = nullptr; size_t used = 0;
  int32_t rc = tfr_decode((tfr_decoder*)dec, p, (size_t)nbytes, 0, isFinal ? 1 : 0, &b, &used);
  if (rc) { throw_for(env, rc, -1); return 0; }
  jlong u = (jlong)used; env->SetLongArrayRegion(consumedOut, 0, 1, &u);
  return (jlong)b;
}
// pipelined form: block t+1 is read from the InputStream into another staging slot while block t is in flight
extern "C" JNIEXPORT jlong JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_decodeSubmit(
    JNIEnv* env, jclass, jlong dec, jobject staged, jlong nbytes, jboolean isFinal) {
  void* p = env->GetDirectBufferAddress(staged);
  tfr_batch* b = nullptr;
  int32_t rc = tfr_decode_submit((tfr_decoder*)dec, p, (size_t)nbytes, 0, isFinal ? 1 : 0, &b);
  if (rc) { throw_for(env, rc, -1); return 0; }
  return (jlong)b;
}
extern "C" JNIEXPORT void JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_batchToHostAsync(JNIEnv* env, jclass, jlong batch) {
  int32_t rc = tfr_batch_to_host_async((tfr_batch*)batch);
  if (rc) { tfr_batch_release((tfr_batch*)batch); throw_for(env, rc, -1); }
}
// where the block after this one starts: known after the batch's frame index, before its rows (the reader submits the next block first)
extern "C" JNIEXPORT jlong JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_batchConsumed(JNIEnv* env, jclass, jlong batch) {
  size_t used = 0;
  int32_t rc = tfr_batch_consumed((tfr_batch*)batch, &used);
  if (rc) { tfr_batch_release((tfr_batch*)batch); throw_for(env, rc, -1); return 0; }
  return (jlong)used;
}
extern "C" JNIEXPORT jlongArray JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_batchStatus(JNIEnv* env, jclass, jlong batch) {
  tfr_batch_info i{};
  int32_t rc0 = tfr_batch_status((tfr_batch*)batch, &i);
  if (rc0) { tfr_batch_release((tfr_batch*)batch); throw_for(env, rc0, -1); return nullptr; }   // a CUDA failure: the batch is gone, the task fails
  jlong v[6] = {i.n_rows, i.n_records, i.consumed_bytes, i.error_code, i.error_row, i.error_field};
  jlongArray a = env->NewLongArray(6); env->SetLongArrayRegion(a, 0, 6, v);
  return a;
}
// The Scala iterator calls this once per column, wraps the buffers in OnHeap/OffHeap column vectors (or an
// ArrowColumnVector over the exported ArrowArray) and, after the last delivered row, throws the exception for
// batchStatus().errorCode -- the same point in the row stream at which the reference's iterator would throw.
extern "C" JNIEXPORT jobjectArray JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_batchColumnHost(
    JNIEnv* env, jclass, jlong batch, jint column, jlongArray meta) {
  tfr_batch* b = (tfr_batch*)batch;
  int32_t n = tfr_batch_num_columns(b);
  std::vector<tfr_column> cols(n);
  int32_t rc = tfr_batch_to_host(b, cols.data(), n);
  if (rc) { tfr_batch_release(b); throw_for(env, rc, -1); return nullptr; }      // nothing of this batch is reachable any more
  const tfr_column& c = cols[column];
  jobjectArray out = env->NewObjectArray(5, env->FindClass("java/nio/ByteBuffer"), nullptr);
  env->SetObjectArrayElement(out, 0, env->NewDirectByteBuffer(c.validity, (c.n_rows + 7) / 8));
  for (int l = 0; l < c.n_levels; ++l) env->SetObjectArrayElement(out, 1 + l, env->NewDirectByteBuffer(c.offsets[l], c.n_offsets[l] * 4));
  env->SetObjectArrayElement(out, 4, env->NewDirectByteBuffer(c.values, c.n_values * (c.value_width ? c.value_width : 1)));
  jlong m[4] = {c.n_rows, c.null_count, c.n_levels, c.n_values};
  env->SetLongArrayRegion(meta, 0, 4, m);
  return out;
}
// Device-resident hand-over (supportBatch = true with GPU column vectors): fills the caller's struct ArrowDeviceArray /
// struct ArrowSchema (addresses of off-heap memory the JVM allocated); the array's release callback drops the batch reference.
extern "C" JNIEXPORT void JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_batchExportArrowDevice(
    JNIEnv* env, jclass, jlong batch, jint column, jlong arrowDeviceArrayAddr, jlong arrowSchemaAddr) {
  int32_t rc = tfr_batch_export_arrow_device((tfr_batch*)batch, column, (void*)arrowDeviceArrayAddr, (void*)arrowSchemaAddr);
  if (rc) throw_for(env, rc, -1);
}
// After the last delivered row of a block the iterator calls this: it throws what the reference's next() would have thrown
// for the first failing record (and releases the batch first: the exception ends the task's use of it).
extern "C" JNIEXPORT void JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_batchThrowIfError(JNIEnv* env, jclass, jlong batch) {
  tfr_batch_info i{};
  int32_t rc = tfr_batch_status((tfr_batch*)batch, &i);
  if (rc == 0 && i.error_code == 0) return;
  tfr_batch_release((tfr_batch*)batch);
  throw_for(env, rc ? rc : i.error_code, rc ? -1 : i.error_row);
}
extern "C" JNIEXPORT void JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_batchRelease(JNIEnv*, jclass, jlong batch) { if (batch) tfr_batch_release((tfr_batch*)batch); }
extern "C" JNIEXPORT void JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_encoderDestroy(JNIEnv*, jclass, jlong enc) { if (enc) tfr_encoder_destroy((tfr_encoder*)enc); }
extern "C" JNIEXPORT jlong JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_encoderCreate(JNIEnv* env, jclass, jlong schema, jint device) {
  tfr_encoder* e = nullptr;
  int32_t rc = tfr_encoder_create((const tfr_schema*)schema, device, 0, &e);
  if (rc) { throw_for(env, rc, -1); return 0; }
  return (jlong)e;
}
// columnStructAddrs: addresses of tfr_column structs the Scala writer filled from its row buffer (off-heap)
extern "C" JNIEXPORT jobject JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_encode(JNIEnv* env, jclass, jlong enc, jlongArray columnStructAddrs, jint n) {
  std::vector<tfr_column> cols(n);
  jlong* a = env->GetLongArrayElements(columnStructAddrs, nullptr);
  for (jint i = 0; i < n; ++i) cols[i] = *(const tfr_column*)a[i];
  env->ReleaseLongArrayElements(columnStructAddrs, a, JNI_ABORT);
  void* dev = nullptr; size_t nb = 0; int64_t err_row = -1;
  int32_t rc = tfr_encode((tfr_encoder*)enc, cols.data(), n, 0, &dev, &nb, &err_row);
  if (rc) { throw_for(env, rc, err_row); return nullptr; }
  void* host = nullptr;
  rc = tfr_encoder_result_host((tfr_encoder*)enc, &host, &nb);
  if (rc) { throw_for(env, rc, -1); return nullptr; }
  return env->NewDirectByteBuffer(host, (jlong)nb);          // outputStream.write(...) of these bytes == the reference file
}

// ---- schema inference: DefaultSource.inferSchema -> TensorFlowInferSchema (M/DefaultSource.scala:31-39,48-70) ----
extern "C" JNIEXPORT jlong JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_inferCreate(JNIEnv* env, jclass, jint recordType, jint device) {
  tfr_infer* h = nullptr;
  int32_t rc = tfr_infer_create(recordType, device, &h);
  if (rc) { throw_for(env, rc, -1); return 0; }
  return (jlong)h;
}
extern "C" JNIEXPORT jlong JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_inferUpdate(JNIEnv* env, jclass, jlong infer, jobject block, jlong nbytes, jboolean isFinal) {
  size_t used = 0;
  int32_t rc = tfr_infer_update_block((tfr_infer*)infer, env->GetDirectBufferAddress(block), (size_t)nbytes, 0, isFinal ? 1 : 0, &used);
  if (rc) { throw_for(env, rc, -1); return 0; }
  return (jlong)used;
}
extern "C" JNIEXPORT jobjectArray JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_inferResult(JNIEnv* env, jclass, jlong infer) {
  int32_t n = 0;
  int32_t rc = tfr_infer_result((tfr_infer*)infer, &n);
  if (rc) { throw_for(env, rc, -1); return nullptr; }
  jobjectArray names = env->NewObjectArray(n, env->FindClass("java/lang/String"), nullptr);
  std::vector<jint> codes(n);
  for (int32_t i = 0; i < n; ++i) {
    const char* nm = nullptr; int32_t len = 0, code = 0;
    tfr_infer_name((tfr_infer*)infer, i, &nm, &len, &code);
    env->SetObjectArrayElement(names, i, env->NewStringUTF(std::string(nm, (size_t)len).c_str()));
    codes[i] = code;
  }
  jintArray jc = env->NewIntArray(n);
  env->SetIntArrayRegion(jc, 0, n, codes.data());
  jobjectArray out = env->NewObjectArray(2, env->FindClass("java/lang/Object"), nullptr);
  env->SetObjectArrayElement(out, 0, names);
  env->SetObjectArrayElement(out, 1, jc);
  return out;
}
extern "C" JNIEXPORT void JNICALL Java_com_linkedin_spark_datasources_tfrecord_TfrGpu_inferDestroy(JNIEnv*, jclass, jlong infer) { if (infer) tfr_infer_destroy((tfr_infer*)infer); }
#endif  // TFR_BUILD_JNI
