/* osc_gemm_check.c — an oscillator bank through the C ABI, every voice on its own output slot, checked at sampled times against
 * the fp64 closed form  sum_p amp min(t/A, 1) exp(-t/tau) sin(2 pi f t / sr + phi)  evaluated here, and timed.
 * FRB_OSC_GEMM=0 / 2 in the environment selects the resonator kernel (K1) / the tensor-core kernel (K1G, csrc/osc_gemm.cuh).
 * build: gcc -O2 -std=gnu99 -Iinclude tools/microbench/osc_gemm_check.c -Llibfriendship_b200/lib -lfriendship_b200 -lm -o build/bin/osc_gemm_check
 * run:   build/bin/osc_gemm_check [voices] [partials] [samples] [points per voice] [block]  */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "friendship_b200.h"

static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static uint64_t sm_state = 1;
static uint64_t next64(void) {
    uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
static double uniform01(void) { return (double)(next64() >> 11) / 9007199254740992.0; }
#define OK(x) do { if ((x) != 0) { printf("FAIL %s: %s\n", #x, frb_last_error(r)); return 1; } } while (0)

int main(int argc, char** argv) {
    const uint32_t nv = argc > 1 ? (uint32_t)atoi(argv[1]) : 8;
    const uint64_t np = argc > 2 ? strtoull(argv[2], 0, 10) : 4096;
    const uint64_t ns = argc > 3 ? strtoull(argv[3], 0, 10) : 40000;
    const int npts = argc > 4 ? atoi(argv[4]) : 200;
    const uint64_t block = argc > 5 ? strtoull(argv[5], 0, 10) : ns;
    const double sr = 48000.0, PI = 3.14159265358979323846;
    frb_config cfg; memset(&cfg, 0, sizeof cfg);
    frb_renderer* r = frb_create(&cfg);
    if (!r) { printf("FAIL frb_create: %s\n", frb_last_error(NULL)); return 1; }
    const uint64_t n = nv * np;
    uint64_t* vo = malloc((nv + 1) * sizeof *vo);
    double* freq = malloc(n * sizeof *freq);
    float *amp = malloc(n * 4), *phase = malloc(n * 4), *attack = malloc(n * 4), *tau = malloc(n * 4);
    double* fs = calloc(nv, sizeof *fs);
    for (uint32_t v = 0; v <= nv; v++) vo[v] = v * np;
    for (uint32_t v = 0; v < nv; v++) {
        const double f0 = 55.0 * pow(2.0, v / 12.0);
        for (uint64_t k = 1; k <= np; k++) {
            double f = f0 * (double)k * (1.0 + (uniform01() * 0.004 - 0.002));
            if (f >= sr / 2) f = fmod(f, sr / 2 * 0.98) + 20.0;
            const uint64_t i = v * np + k - 1;
            freq[i] = f; amp[i] = (float)((v % 3 == 1 ? 1e-3 : v % 3 == 2 ? 40.0 : 1.0) / (double)k);   /* voices at different levels */
            phase[i] = (float)(uniform01() * 6.0);
            attack[i] = (float)(48.0 * (1 + k % 7)); tau[i] = (k % 11 == 0) ? 0.0f : (float)(sr * (0.2 + 2.0 / (double)k));
            fs[v] += fabs(amp[i]);
        }
    }
    frb_oscbank_desc d; memset(&d, 0, sizeof d);
    d.n_voices = nv; d.n_partials = n; d.sample_rate = sr; d.voice_offsets = vo; d.freq_hz = freq; d.amp = amp; d.phase = phase; d.attack = attack; d.tau = tau;
    OK(frb_define_oscbank(r, 7, &d));
    OK(frb_add_node(r, 2, FRB_KIND_OSCBANK, 7));
    for (uint32_t v = 0; v < nv; v++) { frb_edge e; e.from = 2; e.to = 0; e.from_slot = v; e.to_slot = v; OK(frb_add_edge(r, e)); }

    float* out = malloc((size_t)nv * ns * sizeof *out);
    float* blk = malloc((size_t)nv * block * sizeof *blk);
    double best = 1e30;
    for (int rep = 0; rep < 3; rep++) {
        const double t0 = now_ms();
        for (uint64_t c0 = 0; c0 < ns; c0 += block) {
            const uint64_t len = ns - c0 < block ? ns - c0 : block;
            OK(frb_fill_buffer(r, blk, nv, len, c0, NULL, NULL, 0));
            for (uint32_t v = 0; v < nv; v++) memcpy(out + (size_t)v * ns + c0, blk + (size_t)v * len, len * sizeof(float));
        }
        const double dt = now_ms() - t0;
        if (rep > 0 && dt < best) best = dt;
    }
    /* sampled check: the attack region, tile and block boundaries, and random times */
    double worst = 0; uint64_t worst_t = 0; uint32_t worst_v = 0; int bad = 0;
    for (uint32_t v = 0; v < nv; v++) {
        for (int q = 0; q < npts; q++) {
            uint64_t t;
            if (q < 8) t = (uint64_t)q * 47 % ns;
            else if (q < 16) t = (384 - 4 + (uint64_t)(q - 8)) % ns;            /* around the end of the longest ramp */
            else if (q < 24) t = (16384 - 4 + (uint64_t)(q - 16)) % ns;        /* around a tile boundary */
            else if (q < 28) t = ns - 1 - (uint64_t)(q - 24);
            else t = next64() % ns;
            double ref = 0;
            for (uint64_t k = 0; k < np; k++) {
                const uint64_t i = v * np + k;
                const double env = (attack[i] > 0 && t < attack[i] ? t / (double)attack[i] : 1.0) * (tau[i] > 0 ? exp(-(double)t / (double)tau[i]) : 1.0);
                double fr = freq[i] / sr; fr -= floor(fr);
                double turns = fr * (double)t; turns -= floor(turns);
                ref += (double)amp[i] * env * sin(2.0 * PI * turns + (double)phase[i]);
            }
            const double err = fabs((double)out[(size_t)v * ns + t] - ref) / fs[v];
            if (!(err <= 1e-5)) bad++;
            if (!(err <= worst)) { worst = err; worst_t = t; worst_v = v; }
        }
    }
    double sum = 0;
    for (size_t i = 0; i < (size_t)nv * ns; i++) sum += out[i];
    frb_stats st; frb_get_stats(r, &st);
    const char* mode = getenv("FRB_OSC_GEMM");
    printf("{\"case\": \"oscbank vs fp64 closed form\", \"FRB_OSC_GEMM\": \"%s\", \"voices\": %u, \"partials\": %llu, \"samples\": %llu, \"block\": %llu, "
           "\"ms_per_render_best\": %.3f, \"partial_samples_per_s\": %.4e, \"points\": %d, \"max_err_of_full_scale\": %.3e, \"at_voice\": %u, \"at_t\": %llu, "
           "\"points_above_1e-5\": %d, \"checksum\": %.9e, \"kernel_launches\": %llu}\n",
           mode ? mode : "", nv, (unsigned long long)np, (unsigned long long)ns, (unsigned long long)block, best,
           (double)nv * (double)np * (double)ns / (best * 1e-3), npts * (int)nv, worst, worst_v, (unsigned long long)worst_t, bad, sum,
           (unsigned long long)st.kernel_launches);
    frb_destroy(r);
    return bad ? 2 : 0;
}
