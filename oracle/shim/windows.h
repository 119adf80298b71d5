// oracle/shim/windows.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal pthread-backed stand-in for the handful of Win32 names the reference codec core
// (screencap.cpp, ans_contexts.cpp, squad.cpp, sub.cpp, ransmt.h) touches, so that the
// UNMODIFIED reference sources under /root/reference compile on Linux into oracle/_ref/.
// Nothing here is part of the product; it only exists so the real reference can be the oracle.
#ifndef SCPR_ORACLE_SHIM_WINDOWS_H
#define SCPR_ORACLE_SHIM_WINDOWS_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <algorithm>
#include <vector>
#include <string>
#include <stdexcept>

typedef uint8_t BYTE;
typedef uint16_t WORD;
typedef uint32_t DWORD;
typedef int32_t LONG;
typedef int BOOL;
typedef void* LPVOID;
typedef void* HMODULE;
typedef long long __int64;
typedef const char* LPCSTR;

#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define INFINITE 0xFFFFFFFFu
#define WAIT_OBJECT_0 0u
#define WINAPI
#define __forceinline inline __attribute__((always_inline))

// ---- handles: events and threads -------------------------------------------------------
struct ShimHandle {
    int kind;  // 0 = event, 1 = thread
    pthread_mutex_t m;
    pthread_cond_t c;
    bool manual, signaled;
    pthread_t th;
};
typedef ShimHandle* HANDLE;

static inline HANDLE CreateEvent(void*, BOOL manual, BOOL initial, const char*) {
    ShimHandle* h = new ShimHandle();
    h->kind = 0;
    pthread_mutex_init(&h->m, NULL);
    pthread_cond_init(&h->c, NULL);
    h->manual = manual != 0;
    h->signaled = initial != 0;
    return h;
}
static inline BOOL SetEvent(HANDLE h) {
    pthread_mutex_lock(&h->m);
    h->signaled = true;
    pthread_cond_broadcast(&h->c);
    pthread_mutex_unlock(&h->m);
    return TRUE;
}
static inline BOOL ResetEvent(HANDLE h) {
    pthread_mutex_lock(&h->m);
    h->signaled = false;
    pthread_mutex_unlock(&h->m);
    return TRUE;
}
static inline DWORD WaitForSingleObject(HANDLE h, DWORD) {
    if (h->kind == 1) {
        pthread_join(h->th, NULL);
        return WAIT_OBJECT_0;
    }
    pthread_mutex_lock(&h->m);
    while (!h->signaled) pthread_cond_wait(&h->c, &h->m);
    if (!h->manual) h->signaled = false;
    pthread_mutex_unlock(&h->m);
    return WAIT_OBJECT_0;
}
static inline DWORD WaitForMultipleObjects(DWORD n, HANDLE* hs, BOOL /*all*/, DWORD t) {
    for (DWORD i = 0; i < n; i++) WaitForSingleObject(hs[i], t);
    return WAIT_OBJECT_0;
}
static inline DWORD SignalObjectAndWait(HANDLE a, HANDLE b, DWORD t, BOOL) {
    SetEvent(a);
    return WaitForSingleObject(b, t);
}
static inline BOOL CloseHandle(HANDLE h) {
    if (!h) return TRUE;
    if (h->kind == 0) {
        pthread_mutex_destroy(&h->m);
        pthread_cond_destroy(&h->c);
    }
    delete h;
    return TRUE;
}

typedef DWORD (*LPTHREAD_START_ROUTINE)(LPVOID);
struct ShimThreadStart {
    LPTHREAD_START_ROUTINE fn;
    LPVOID arg;
};
static inline void* shim_thread_tramp(void* p) {
    ShimThreadStart s = *(ShimThreadStart*)p;
    delete (ShimThreadStart*)p;
    s.fn(s.arg);
    return NULL;
}
static inline HANDLE CreateThread(void*, size_t stack, LPTHREAD_START_ROUTINE fn, LPVOID arg, DWORD, DWORD* tid) {
    ShimHandle* h = new ShimHandle();
    h->kind = 1;
    pthread_attr_t at;
    pthread_attr_init(&at);
    // ransmt.h:116-121 puts a 256 KiB buffer on the worker's stack
    if (stack < (size_t)1 << 20) stack = (size_t)1 << 20;
    pthread_attr_setstacksize(&at, stack);
    ShimThreadStart* s = new ShimThreadStart{fn, arg};
    pthread_create(&h->th, &at, shim_thread_tramp, s);
    pthread_attr_destroy(&at);
    if (tid) *tid = 1;
    return h;
}

// ---- critical sections ------------------------------------------------------------------
typedef pthread_mutex_t CRITICAL_SECTION;
static inline void InitializeCriticalSection(CRITICAL_SECTION* cs) {
    pthread_mutexattr_t a;
    pthread_mutexattr_init(&a);
    pthread_mutexattr_settype(&a, PTHREAD_MUTEX_RECURSIVE);
    pthread_mutex_init(cs, &a);
    pthread_mutexattr_destroy(&a);
}
static inline void DeleteCriticalSection(CRITICAL_SECTION* cs) { pthread_mutex_destroy(cs); }
static inline void EnterCriticalSection(CRITICAL_SECTION* cs) { pthread_mutex_lock(cs); }
static inline void LeaveCriticalSection(CRITICAL_SECTION* cs) { pthread_mutex_unlock(cs); }

// ---- processor count (screencap.cpp:1459-1461) -----------------------------------------------
// The harness sets this; 1 gives the only deterministic (canonical) bitstream.
extern int g_shim_nproc;
struct SYSTEM_INFO {
    DWORD dwNumberOfProcessors;
};
static inline void GetSystemInfo(SYSTEM_INFO* si) { si->dwNumberOfProcessors = g_shim_nproc > 0 ? g_shim_nproc : 1; }

// ---- min/max macros, defined AFTER the C++ headers (screencap.cpp:79 mixes uint/int) -----
#ifndef min
#define min(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef max
#define max(a, b) (((a) > (b)) ? (a) : (b))
#endif


// ---- high-resolution counter (only the reference's own TIMING mode uses it, screencap.cpp:83-85, 325-341, 1096-1268) ------
#include <time.h>
union LARGE_INTEGER {
    long long QuadPart;
};
static inline BOOL QueryPerformanceFrequency(LARGE_INTEGER* f) {
    f->QuadPart = 1000000000LL;
    return TRUE;
}
static inline BOOL QueryPerformanceCounter(LARGE_INTEGER* c) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    c->QuadPart = (long long)ts.tv_sec * 1000000000LL + ts.tv_nsec;
    return TRUE;
}

#endif
