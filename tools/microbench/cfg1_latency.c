/* cfg1_latency.c — the cfg1 graph (BASELINE.json configs[0]) driven from plain C through frb_fill_buffer with HOST
 * buffers, streamed in 512-sample calls: what one call costs without a Python binding in the way.
 * build: gcc -O2 -std=gnu99 -Iinclude tools/microbench/cfg1_latency.c -Llibfriendship_b200/lib -lfriendship_b200 -lm -o /tmp/cfg1_latency
 * run:   LD_LIBRARY_PATH=libfriendship_b200/lib /tmp/cfg1_latency [block] [calls] */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "friendship_b200.h"

static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static double now_us(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }
static int cmp(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }

#define OK(x) do { if ((x) != 0) { printf("FAIL %s: %s\n", #x, frb_last_error(r)); return 1; } } while (0)
static int edge(frb_renderer* r, uint32_t from, uint32_t to, uint32_t fs, uint32_t ts) {
    frb_edge e; e.from = from; e.to = to; e.from_slot = fs; e.to_slot = ts; return frb_add_edge(r, e);
}

int main(int argc, char** argv) {
    const uint64_t block = argc > 1 ? strtoull(argv[1], 0, 10) : 512;
    const int calls = argc > 2 ? atoi(argv[2]) : 2000;
    frb_config cfg; memset(&cfg, 0, sizeof cfg);
    if (getenv("CFG1_FLAGS")) cfg.flags = (uint32_t)atoi(getenv("CFG1_FLAGS"));   /* e.g. 2 = FRB_FLAG_NO_JIT, 4 = FRB_FLAG_JIT_EAGER */
    frb_renderer* r = frb_create(&cfg);
    if (!r) { printf("FAIL frb_create: %s\n", frb_last_error(NULL)); return 1; }
    /* handles: 1 = constants, 2 gain, 3 delay, 4 wet gain, 5 sum, 6 min, 7 mod, 8 div (tests/graphs.py build_cfg1_graph) */
    OK(frb_add_node(r, 1, FRB_KIND_F32CONSTANT, 0));
    OK(frb_add_node(r, 2, FRB_KIND_MULTIPLY, 0)); OK(edge(r, 0, 2, 0, 0)); OK(edge(r, 1, 2, bits(0.5f), 1));
    OK(frb_add_node(r, 3, FRB_KIND_DELAY, 0));    OK(edge(r, 2, 3, 0, 0)); OK(edge(r, 1, 3, bits(12000.0f), 1));
    OK(frb_add_node(r, 4, FRB_KIND_MULTIPLY, 0)); OK(edge(r, 3, 4, 0, 0)); OK(edge(r, 1, 4, bits(0.35f), 1));
    OK(frb_add_node(r, 5, FRB_KIND_SUM2, 0));     OK(edge(r, 2, 5, 0, 0)); OK(edge(r, 4, 5, 0, 1)); OK(edge(r, 5, 0, 0, 0));
    OK(frb_add_node(r, 6, FRB_KIND_MINIMUM, 0));  OK(edge(r, 0, 6, 0, 0)); OK(edge(r, 1, 6, bits(0.25f), 1));
    OK(frb_add_node(r, 7, FRB_KIND_MODULO, 0));   OK(edge(r, 6, 7, 0, 0)); OK(edge(r, 1, 7, bits(0.1f), 1));
    OK(frb_add_node(r, 8, FRB_KIND_DIVIDE, 0));   OK(edge(r, 7, 8, 0, 0)); OK(edge(r, 1, 8, bits(3.0f), 1)); OK(edge(r, 8, 0, 0, 1));

    float* x = malloc(block * sizeof(float));
    float* out = malloc(2 * block * sizeof(float));
    double* us = malloc(calls * sizeof(double));
    uint64_t offs[2]; offs[0] = 0; offs[1] = block;
    uint64_t idx = 0;
    double checksum = 0;
    for (int c = -200; c < calls; c++) {                 /* 200 warm-up calls: the stage JIT compiles beside the loop */
        for (uint64_t i = 0; i < block; i++) x[i] = (float)sin(2.0 * M_PI * 440.0 * (double)(idx + i) / 48000.0);
        const double t0 = now_us();
        OK(frb_fill_buffer(r, out, 2, block, idx, x, offs, 1));
        const double t1 = now_us();
        if (c >= 0) us[c] = t1 - t0;
        checksum += out[0] + out[block];
        idx += block;
    }
    /* medians of the four quarters of the run, in call order: shows a tier change (interpreter -> compiled stage) */
    double quarter[4];
    for (int k = 0; k < 4; k++) {
        const int a = calls * k / 4, n = calls * (k + 1) / 4 - a;
        double* tmp = malloc((n > 0 ? n : 1) * sizeof(double));
        memcpy(tmp, us + a, n * sizeof(double));
        qsort(tmp, n, sizeof(double), cmp);
        quarter[k] = n ? tmp[n / 2] : 0;
        free(tmp);
    }
    qsort(us, calls, sizeof(double), cmp);
    printf("{\"case\": \"cfg1 through the C ABI, host buffers\", \"block\": %llu, \"calls\": %d, \"us_per_call_median\": %.2f, "
           "\"us_p10\": %.2f, \"us_p90\": %.2f, \"us_median_by_quarter\": [%.2f, %.2f, %.2f, %.2f], \"realtime_factor\": %.1f, \"checksum\": %.6f}\n",
           (unsigned long long)block, calls, us[calls / 2], us[calls / 10], us[calls * 9 / 10], quarter[0], quarter[1], quarter[2], quarter[3],
           (double)block / 48000.0 * 1e6 / us[calls / 2], checksum);
    frb_stats st; frb_get_stats(r, &st);
    printf("{\"stats\": {\"kernel_launches\": %llu, \"jit_launches\": %llu, \"interp_launches\": %llu, \"h2d_bytes\": %llu, \"d2h_bytes\": %llu, \"schedule_builds\": %llu}}\n",
           (unsigned long long)st.kernel_launches, (unsigned long long)st.jit_launches, (unsigned long long)st.interp_launches,
           (unsigned long long)st.h2d_bytes, (unsigned long long)st.d2h_bytes, (unsigned long long)st.schedule_builds);
    frb_destroy(r);
    return 0;
}
