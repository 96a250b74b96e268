/* abi_smoke.c — a plain-C99 client of include/friendship_b200.h and friendship_dispatch.h: proves the headers compile
 * as C and that a non-Python host can drive the renderer.  argv[1] = CUDA device (-1: planning only, no rendering).
 * Replays the reference's render_delay test (tests/render_prim.rs:101-129) through Dispatch when a device is given. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "friendship_b200.h"
#include "friendship_dispatch.h"

static float g_buf[4];
static int g_calls;
static void on_audio(void* user, const float* buffer, uint32_t n_slots, uint64_t n_times, uint64_t idx) {
    (void)user; (void)idx;
    if (n_slots == 1 && n_times == 4) memcpy(g_buf, buffer, sizeof g_buf);
    g_calls++;
}

#define CHECK(x) do { int rc_ = (x); if (rc_ != 0) { printf("FAIL %s -> %d (%s)\n", #x, rc_, frd_last_error(d)); return 1; } } while (0)

int main(int argc, char** argv) {
    frb_config cfg;
    frd_client client;
    frd_dispatch* d;
    frb_edge e;
    int device = argc > 1 ? atoi(argv[1]) : -1;
    memset(&cfg, 0, sizeof cfg);
    memset(&client, 0, sizeof client);
    cfg.device = device;
    client.audio_rendered = on_audio;
    d = frd_create(&cfg, &client);
    if (!d) { printf("FAIL frd_create: %s\n", frb_last_error(NULL)); return 1; }
    CHECK(frd_add_node(d, 1, "{\"name\":\"Delay\",\"sha256\":null,\"urls\":[\"primitive:///Delay\"]}"));
    e.from = 1; e.to = 0; e.from_slot = 0; e.to_slot = 0;
    CHECK(frd_add_edge(d, e));
    CHECK(frd_add_node(d, 2, "{\"name\":\"F32Constant\",\"sha256\":null,\"urls\":[\"primitive:///F32Constant\"]}"));
    e.from = 2; e.to = 1; e.from_slot = 0x3f000000u; e.to_slot = 0;       /* 0.5f */
    CHECK(frd_add_edge(d, e));
    e.from_slot = 0x40000000u; e.to_slot = 1;                             /* 2.0f frames */
    CHECK(frd_add_edge(d, e));
    e.to_slot = 0;
    if (frd_add_edge(d, e) != FRD_E_SLOT_ALREADY_CONNECTED) { printf("FAIL expected SlotAlreadyConnected\n"); return 1; }
    {
        uint32_t words[256];
        int64_t n = frb_dump_schedule(frd_renderer(d), 1, words, 256);
        if (n <= 8 || words[0] != 0x53425246u) { printf("FAIL dump_schedule %lld\n", (long long)n); return 1; }
    }
    if (device >= 0) {
        CHECK(frd_render_range(d, 0, 4, 1, NULL, NULL, 0));
        if (g_calls != 1 || g_buf[0] != 0.0f || g_buf[1] != 0.0f || g_buf[2] != 0.5f || g_buf[3] != 0.5f) {
            printf("FAIL render_delay: %g %g %g %g\n", g_buf[0], g_buf[1], g_buf[2], g_buf[3]);
            return 1;
        }
    } else if (frd_render_range(d, 0, 4, 1, NULL, NULL, 0) != FRB_E_NO_DEVICE) {
        printf("FAIL planning-only handle rendered\n");
        return 1;
    }
    frd_destroy(d);
    printf("abi_smoke ok (device %d, %s)\n", device, frb_version());
    return 0;
}
