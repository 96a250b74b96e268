/* cfg4_multi.c — BASELINE.json configs[3] (65,536 partials x 64 voices, per-voice Delay/mix, 48 kHz x 10 s) rendered from
 * plain C through ONE renderer handle on N B200s: frb_config.n_devices = N puts the voice sharding and the exchange
 * behind the C ABI (csrc/multi.cu), which is how a Rust `Dispatch<B200Renderer, C>` reaches several GPUs
 * (reference src/dispatch.rs:99-106: the caller owns one renderer).
 * build: gcc -O2 -std=gnu99 -Iinclude tools/microbench/cfg4_multi.c -Llibfriendship_b200/lib -lfriendship_b200 -lm -o /tmp/cfg4_multi
 * run:   LD_LIBRARY_PATH=libfriendship_b200/lib /tmp/cfg4_multi [n_devices] [steps] [voices] [partials] [samples]
 * The detune sequence is this program's own (splitmix64), not numpy's PCG64: same shape and statistics as the bench's bank. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "friendship_b200.h"

static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static uint64_t sm_state = 1;
static double uniform01(void) {
    uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (double)(z >> 11) / 9007199254740992.0;
}
#define OK(x) do { if ((x) != 0) { printf("FAIL %s: %s\n", #x, frb_last_error(r)); return 1; } } while (0)
static int edge(frb_renderer* r, uint32_t from, uint32_t to, uint32_t fs, uint32_t ts) {
    frb_edge e; e.from = from; e.to = to; e.from_slot = fs; e.to_slot = ts; return frb_add_edge(r, e);
}

int main(int argc, char** argv) {
    const uint32_t n_dev = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    const int steps = argc > 2 ? atoi(argv[2]) : 3;
    const uint32_t nv = argc > 3 ? (uint32_t)atoi(argv[3]) : 64;
    const uint64_t np = argc > 4 ? strtoull(argv[4], 0, 10) : 65536;
    const uint64_t ns = argc > 5 ? strtoull(argv[5], 0, 10) : 480000;
    const double sr = 48000.0;
    frb_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.n_devices = n_dev;
    frb_renderer* r = frb_create(&cfg);
    if (!r) { printf("FAIL frb_create: %s\n", frb_last_error(NULL)); return 1; }

    const uint64_t n = nv * np;
    uint64_t* vo = malloc((nv + 1) * sizeof *vo);
    double* freq = malloc(n * sizeof *freq);
    float *amp = malloc(n * 4), *phase = calloc(n, 4), *attack = malloc(n * 4), *tau = malloc(n * 4);
    for (uint32_t v = 0; v <= nv; v++) vo[v] = v * np;
    for (uint32_t v = 0; v < nv; v++) {
        const double f0 = 55.0 * pow(2.0, v / 12.0);
        for (uint64_t k = 1; k <= np; k++) {
            double f = f0 * (double)k * (1.0 + (uniform01() * 0.004 - 0.002));
            if (f >= sr / 2) f = fmod(f, sr / 2 * 0.98) + 20.0;
            const uint64_t i = v * np + k - 1;
            freq[i] = f; amp[i] = (float)(1.0 / (double)k);
            attack[i] = (float)(48.0 * (1 + k % 7)); tau[i] = (float)(sr * (0.2 + 2.0 / (double)k));
        }
    }
    frb_oscbank_desc d; memset(&d, 0, sizeof d);
    d.n_voices = nv; d.n_partials = n; d.sample_rate = sr; d.voice_offsets = vo; d.freq_hz = freq; d.amp = amp; d.phase = phase; d.attack = attack; d.tau = tau;
    OK(frb_define_oscbank(r, 7, &d));
    /* graph of workloads/banks.py build_voice_mix_graph: handle 1 constants, 2 the bank, then per voice Delay, Multiply, Sum2 (+ Sum2 into the mix) */
    OK(frb_add_node(r, 1, FRB_KIND_F32CONSTANT, 0));
    OK(frb_add_node(r, 2, FRB_KIND_OSCBANK, 7));
    uint32_t h = 3, total = 0;
    for (uint32_t v = 0; v < nv; v++) {
        const uint32_t dl = h++, wet = h++, voice = h++;
        OK(frb_add_node(r, dl, FRB_KIND_DELAY, 0));      OK(edge(r, 2, dl, v, 0));   OK(edge(r, 1, dl, bits(4800.0f + 37.0f * v), 1));
        OK(frb_add_node(r, wet, FRB_KIND_MULTIPLY, 0));  OK(edge(r, dl, wet, 0, 0)); OK(edge(r, 1, wet, bits(0.3f), 1));
        OK(frb_add_node(r, voice, FRB_KIND_SUM2, 0));    OK(edge(r, 2, voice, v, 0)); OK(edge(r, wet, voice, 0, 1));
        if (!total) total = voice;
        else { const uint32_t t = h++; OK(frb_add_node(r, t, FRB_KIND_SUM2, 0)); OK(edge(r, total, t, 0, 0)); OK(edge(r, voice, t, 0, 1)); total = t; }
    }
    OK(edge(r, total, 0, 0, 0));

    float* out = malloc(ns * sizeof *out);
    double best = 1e30, first = 0;
    for (int s = -2; s < steps; s++) {                     /* two warm-up renders: schedules, stage JIT, buffers */
        const double t0 = now_ms();
        OK(frb_fill_buffer(r, out, 1, ns, 0, NULL, NULL, 0));
        const double dt = now_ms() - t0;
        if (s == -2) first = dt;
        if (s >= 0 && dt < best) best = dt;
    }
    double peak = 0, sum = 0;
    for (uint64_t i = 0; i < ns; i++) { if (fabs(out[i]) > peak) peak = fabs(out[i]); sum += out[i]; }
    frb_stats st; frb_get_stats(r, &st);
    printf("{\"case\": \"cfg4 through the C ABI, one handle, host output buffer\", \"n_devices\": %u, \"voices\": %u, \"partials\": %llu, \"samples\": %llu, "
           "\"ms_per_render_best\": %.3f, \"partial_samples_per_s\": %.4e, \"first_call_ms\": %.1f, \"peak\": %.4f, \"checksum\": %.6f, \"kernel_launches\": %llu, \"jit_launches\": %llu, \"interp_launches\": %llu}\n",
           n_dev, nv, (unsigned long long)np, (unsigned long long)ns, best, (double)nv * (double)np * (double)ns / (best * 1e-3), first, peak, sum,
           (unsigned long long)st.kernel_launches, (unsigned long long)st.jit_launches, (unsigned long long)st.interp_launches);
    frb_destroy(r);
    return 0;
}
