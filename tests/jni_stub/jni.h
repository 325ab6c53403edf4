// minimal stand-in for <jni.h>: only for a -fsyntax-only check of the shim in an image without a JDK
#pragma once
#include <cstdint>
#define JNIEXPORT
#define JNICALL
#define JNI_ABORT 2
typedef int32_t jint; typedef int64_t jlong; typedef uint8_t jboolean; typedef int8_t jbyte; typedef int32_t jsize;
struct _jobject {}; typedef _jobject* jobject; typedef jobject jclass; typedef jobject jstring; typedef jobject jarray;
typedef jarray jobjectArray; typedef jarray jintArray; typedef jarray jlongArray; typedef jarray jbooleanArray; typedef jarray jbyteArray;
struct JNIEnv {
  jclass FindClass(const char*); jint ThrowNew(jclass, const char*); jsize GetArrayLength(jarray);
  jboolean* GetBooleanArrayElements(jbooleanArray, jboolean*); void ReleaseBooleanArrayElements(jbooleanArray, jboolean*, jint);
  jint* GetIntArrayElements(jintArray, jboolean*); void ReleaseIntArrayElements(jintArray, jint*, jint);
  jlong* GetLongArrayElements(jlongArray, jboolean*); void ReleaseLongArrayElements(jlongArray, jlong*, jint);
  jobject GetObjectArrayElement(jobjectArray, jsize); void SetObjectArrayElement(jobjectArray, jsize, jobject);
  const char* GetStringUTFChars(jstring, jboolean*); void ReleaseStringUTFChars(jstring, const char*);
  void* GetDirectBufferAddress(jobject); jobject NewDirectByteBuffer(void*, jlong);
  jstring NewStringUTF(const char*); jintArray NewIntArray(jsize); void SetIntArrayRegion(jintArray, jsize, jsize, const jint*);
  jlongArray NewLongArray(jsize); jobjectArray NewObjectArray(jsize, jclass, jobject); void SetLongArrayRegion(jlongArray, jsize, jsize, const jlong*);
};
