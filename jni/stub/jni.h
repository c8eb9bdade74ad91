/*
 * Declarations-only stand-in for the JDK's <jni.h>, so that jni/biospectra_gpu_jni.c can be compile-checked in an
 * image without a JDK (tests/test_abi.py: gcc -fsyntax-only -Wall -Werror).  It declares the JNI types and exactly the
 * JNIEnv functions the shim calls, with the signatures of the JNI specification (Java SE 7, chapter 4); the slot
 * order is NOT the real table's, so nothing built against this header may ever be loaded into a JVM.
 */
#ifndef BSP_STUB_JNI_H
#define BSP_STUB_JNI_H

#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_COMMIT 1

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef uint16_t jchar;
typedef int16_t jshort;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;

struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jthrowable;
typedef jobject jarray;
typedef jarray jobjectArray;
typedef jarray jbyteArray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jfloatArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
    jclass (JNICALL *FindClass)(JNIEnv* env, const char* name);
    jint (JNICALL *ThrowNew)(JNIEnv* env, jclass clazz, const char* msg);
    void (JNICALL *DeleteLocalRef)(JNIEnv* env, jobject obj);
    jstring (JNICALL *NewStringUTF)(JNIEnv* env, const char* utf);
    const char* (JNICALL *GetStringUTFChars)(JNIEnv* env, jstring str, jboolean* isCopy);
    void (JNICALL *ReleaseStringUTFChars)(JNIEnv* env, jstring str, const char* chars);
    jsize (JNICALL *GetArrayLength)(JNIEnv* env, jarray array);
    jobjectArray (JNICALL *NewObjectArray)(JNIEnv* env, jsize len, jclass clazz, jobject init);
    jobject (JNICALL *GetObjectArrayElement)(JNIEnv* env, jobjectArray array, jsize index);
    void (JNICALL *SetObjectArrayElement)(JNIEnv* env, jobjectArray array, jsize index, jobject val);
    jlongArray (JNICALL *NewLongArray)(JNIEnv* env, jsize len);
    jbyte* (JNICALL *GetByteArrayElements)(JNIEnv* env, jbyteArray array, jboolean* isCopy);
    jint* (JNICALL *GetIntArrayElements)(JNIEnv* env, jintArray array, jboolean* isCopy);
    jfloat* (JNICALL *GetFloatArrayElements)(JNIEnv* env, jfloatArray array, jboolean* isCopy);
    void (JNICALL *ReleaseByteArrayElements)(JNIEnv* env, jbyteArray array, jbyte* elems, jint mode);
    void (JNICALL *ReleaseIntArrayElements)(JNIEnv* env, jintArray array, jint* elems, jint mode);
    void (JNICALL *ReleaseFloatArrayElements)(JNIEnv* env, jfloatArray array, jfloat* elems, jint mode);
    void (JNICALL *GetByteArrayRegion)(JNIEnv* env, jbyteArray array, jsize start, jsize len, jbyte* buf);
    void (JNICALL *SetIntArrayRegion)(JNIEnv* env, jintArray array, jsize start, jsize len, const jint* buf);
    void (JNICALL *SetLongArrayRegion)(JNIEnv* env, jlongArray array, jsize start, jsize len, const jlong* buf);
    void (JNICALL *SetFloatArrayRegion)(JNIEnv* env, jfloatArray array, jsize start, jsize len, const jfloat* buf);
};

#endif
