/* Trace driver for the reference's libcachemanager.so (TEST INFRASTRUCTURE ONLY, see
 * oracle/__init__.py).  Replays a pre-marshalled int32 trace through ev_lookup
 * (/root/reference/mixed_precs_caching/cache_manager.cpp:231) one sample at a time, exactly
 * as cache_algo/cpp_socket_client.py:119 calls it, without the Python call overhead, and
 * returns the elapsed wall-clock seconds.  Our own code; compiled into oracle/_ref/. */
#include <stddef.h>
#include <string.h>
#include <time.h>

typedef float *(*ev_lookup_fn)(int *);

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* trace: n samples of n_tables int32 row ids (sample-major).  If out != NULL the n_tables*dim
 * floats of every answer are copied to out[i] (the caller must copy the library-owned buffer,
 * cache_manager.hpp:51).  Returns seconds; *checksum accumulates the first float of each answer
 * so the loop cannot be optimised away. */
double ref_drive(ev_lookup_fn fn, const int *trace, long n, int n_tables, int dim, float *out, double *checksum) {
    int ids[64];
    double acc = 0.0;
    const size_t row = (size_t)n_tables * (size_t)dim;
    const double t0 = now_s();
    for (long i = 0; i < n; ++i) {
        memcpy(ids, trace + (size_t)i * n_tables, sizeof(int) * (size_t)n_tables);
        const float *r = fn(ids);
        if (r == NULL) return -1.0;
        if (out != NULL) memcpy(out + (size_t)i * row, r, sizeof(float) * row);
        acc += r[0];
    }
    const double t1 = now_s();
    if (checksum) *checksum = acc;
    return t1 - t0;
}
